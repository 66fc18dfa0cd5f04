"""CPU model (numpy + Python integers) of the MODULAR int8 emulation of an fp64 NT product -- "Ozaki scheme II":
C = A B^T with A [M,K], B [N,K] fp64, computed from 16 int8 x int8 -> int32 GEMMs (one per modulus) and a Chinese-remainder
reconstruction, instead of the 36 digit-plane products of the scheme battgp_b200/csrc/ozaki.cu implements today.

This file is design groundwork for the next kernel generation ("Known gaps" (1)); nothing in the product
imports it.  It fixes, in executable form, every arithmetic decision the CUDA version has to reproduce bit for bit:

  * scaling: per row a power of two 2^e_i >= 2 * max|a_i.| ; a'_ik = trunc(a_ik * 2^(BETA - e_i)), |a'| < 2^(BETA-1)
    (same row-scale rule as oz_slice_kernel, BETA = 55 fractional bits instead of 8 x 7-bit digits)
  * residues: r = a' mod p in the symmetric range [-p/2, p/2) -> int8 for the 16 pairwise-coprime moduli MODULI (all <= 256)
  * products: G_j = A_j B_j^T in int32 (|G_j| <= K * 2^14: K <= 2^17)
  * reconstruction: t_j = G_j mod p_j in [0, p_j);  S = sum_j W_j t_j with W_j = (P/p_j) * ((P/p_j)^-1 mod p_j) mod P held
    as four 32-bit limbs (every limb sum < 2^44: one IMAD.WIDE per limb and modulus);  q = rint(S / P) from an fp64
    estimate (safe: |C'| < P / 8 keeps frac(S/P) within 1/8 of an integer);  C' = S - q P exactly in limbs -> signed
    128-bit -> double with one rounding;  C = C' * 2^(e_i + e_j - 2 BETA).
"""
from __future__ import annotations

import math

import numpy as np

BETA = 55
MODULI = (256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193)
P = math.prod(MODULI)
W = tuple(((P // p) * pow(P // p, -1, p)) % P for p in MODULI)
W_LIMBS = np.array([[(w >> (32 * k)) & 0xFFFFFFFF for k in range(4)] for w in W], dtype=np.uint64)   # [16, 4]
P_LIMBS = np.array([(P >> (32 * k)) & 0xFFFFFFFF for k in range(4)], dtype=np.uint64)
INV_P = 1.0 / float(P)


def check_constants(k_max: int = 2048) -> None:
    for i, p in enumerate(MODULI):
        for q in MODULI[i + 1:]:
            assert math.gcd(p, q) == 1, (p, q)
    assert P.bit_length() <= 128
    # |C'| <= K * 2^(2 (BETA-1)) must stay below P / 8 (margin used by the fp64 estimate of q)
    assert k_max * (1 << (2 * (BETA - 1))) < P // 8, "not enough moduli for this K"


def row_scale_exponent(a: np.ndarray) -> np.ndarray:
    """e_i with 2^e_i >= 2 max_k |a_ik| (e_i = exponent of the row maximum + 2 in frexp terms; zero rows get 0)."""
    m = np.max(np.abs(a), axis=1)
    _, e = np.frexp(m)                   # m = f * 2^e, 0.5 <= f < 1  ->  m < 2^e
    return np.where(m > 0, e + 1, 0).astype(np.int64)


def to_scaled_int(a: np.ndarray, e: np.ndarray) -> np.ndarray:
    """a' = trunc(a * 2^(BETA - e_i)) as int64 (|a'| < 2^(BETA-1) <= 2^54: exact in fp64 before the cast)."""
    return np.trunc(np.ldexp(a, (BETA - e)[:, None])).astype(np.int64)


def residues(ai: np.ndarray) -> np.ndarray:
    """[16, rows, K] int8, symmetric residues."""
    out = np.empty((len(MODULI),) + ai.shape, dtype=np.int8)
    for j, p in enumerate(MODULI):
        r = np.mod(ai, p)                                # [0, p)
        r = np.where(r >= (p + 1) // 2, r - p, r)        # [-p/2, p/2)
        out[j] = r.astype(np.int8)
    return out


def int8_products(ra: np.ndarray, rb: np.ndarray) -> np.ndarray:
    """G_j = A_j B_j^T with int32 accumulation (what tcgen05.mma kind::i8 produces)."""
    g = np.einsum("jmk,jnk->jmn", ra.astype(np.int32), rb.astype(np.int32), dtype=np.int64)
    assert np.abs(g).max() < 2 ** 31
    return g.astype(np.int32)


def crt_reconstruct(g: np.ndarray) -> np.ndarray:
    """Exact C' (as float64 after ONE rounding) from the 16 int32 accumulators, using only operations the GPU epilogue
    has: int32 remainder by a constant, 32x32->64-bit multiply-add, fp64 estimate of the quotient, 64-bit carries."""
    m, n = g.shape[1:]
    s = np.zeros((4, m, n), dtype=np.uint64)
    est = np.zeros((m, n))
    for j, p in enumerate(MODULI):
        t = np.mod(g[j].astype(np.int64), p).astype(np.uint64)           # [0, p)
        for k in range(4):
            s[k] += W_LIMBS[j, k] * t                                    # < 2^32 * 2^8 * 16 = 2^44
        est += t.astype(np.float64) * (float(W[j]) * INV_P)              # S / P, relative error ~1e-15 * 16
    q = np.rint(est).astype(np.int64)                                    # 0 <= q < 16 * 256
    r = s.astype(np.int64) - q[None] * P_LIMBS.astype(np.int64)[:, None, None]     # |.| < 2^45 per limb
    # carry-normalise to limbs in [0, 2^32) below a signed top limb
    for k in range(3):
        c = r[k] >> 32                                                   # floor division (arithmetic shift)
        r[k] -= c << 32
        r[k + 1] += c
    hi = (r[3] << 32) + r[2]                                             # signed, |hi| < 2^59
    lo = ((r[1] << 32) + r[0]).astype(np.uint64)                         # [0, 2^64)
    neg = hi < 0
    # two's-complement negate (hi, lo) where negative, so that both halves are non-negative magnitudes
    lo_n = (~lo + np.uint64(1))
    hi_n = ~hi + (lo == 0)
    hi_m = np.where(neg, hi_n, hi).astype(np.uint64)
    lo_m = np.where(neg, lo_n, lo)
    mag = hi_m.astype(np.float64) * 2.0 ** 64 + lo_m.astype(np.float64)  # <= 2 roundings of 2^-53 relative
    return np.where(neg, -mag, mag)


def gemm_nt_modular(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ea, eb = row_scale_exponent(a), row_scale_exponent(b)
    g = int8_products(residues(to_scaled_int(a, ea)), residues(to_scaled_int(b, eb)))
    cp = crt_reconstruct(g)
    return np.ldexp(cp, (ea[:, None] + eb[None, :] - 2 * BETA))


def exact_scaled_product(a: np.ndarray, b: np.ndarray):
    """Python-integer reference of C' (no rounding anywhere) for the tests."""
    ea, eb = row_scale_exponent(a), row_scale_exponent(b)
    ai, bi = to_scaled_int(a, ea), to_scaled_int(b, eb)
    c = [[sum(int(x) * int(y) for x, y in zip(ai[i], bi[j])) for j in range(b.shape[0])] for i in range(a.shape[0])]
    return c, ea, eb
