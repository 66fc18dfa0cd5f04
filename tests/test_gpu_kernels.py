"""GPU parity tests of the individual C-ABI entry points against the CPU oracle (oracle/gp_oracle.py) on seeded
inputs.  Tolerances are stated per test; all arithmetic is fp64."""
import math

import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _t(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


def _specs():
    from battgp_b200 import engine as E
    return {
        "battgp": (E.battgp_spec(), orc.battgp_spec()),
        "rbf_iso": (E.scaled_rbf_spec(3, 3.0, 2.0), orc.scaled_rbf_spec(3, 3.0, 2.0)),
        "matern_periodic": (E.matern_periodic_spec(), orc.matern_periodic_spec()),
    }


# ----------------------------------------------------------------------------------------------- gemm_nt
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 64, 16), (128, 128, 128), (300, 128, 128),
                                   (257, 391, 70), (1000, 1000, 64), (2048, 1536, 512), (130, 2500, 1031)])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_gemm_nt_matches_fp64_matmul(eng, M, N, K, beta):
    g = torch.Generator(device="cpu").manual_seed(M * 7919 + N * 31 + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    B = torch.randn(N, K, generator=g, dtype=torch.float64)
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64)
    ref = -1.5 * A @ B.T + beta * C0
    from battgp_b200.engine import alloc_matrix
    Ad, Bd, Cd = alloc_matrix(M, K, DEV), alloc_matrix(N, K, DEV), alloc_matrix(M, N, DEV)
    Ad.copy_(A); Bd.copy_(B); Cd.copy_(C0)
    eng.gemm_nt(Ad, Bd, Cd, alpha=-1.5, beta=beta)
    err = (Cd.cpu() - ref).abs().max().item()
    assert err <= 1e-12 * max(1.0, K) , err


def test_gemm_nt_unaligned_leading_dims(eng):
    # odd leading dimensions force the 8-byte cp.async path and scalar epilogue
    M, N, K = 131, 77, 45
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    B = torch.randn(N, K, generator=g, dtype=torch.float64)
    Ad = torch.empty(M, K + 1, dtype=torch.float64, device=DEV)[:, 1:] if False else torch.empty(M, 47, dtype=torch.float64, device=DEV)[:, :K]
    Bd = torch.empty(N, 49, dtype=torch.float64, device=DEV)[:, :K]
    Cd = torch.empty(M, 79, dtype=torch.float64, device=DEV)[:, :N]
    Ad.copy_(A); Bd.copy_(B); Cd.zero_()
    eng.gemm_nt(Ad, Bd, Cd)
    assert (Cd.cpu() - A @ B.T).abs().max().item() < 1e-11


def test_gemm_nt_tri_mask_leaves_upper_untouched(eng):
    n, k = 700, 200
    g = torch.Generator().manual_seed(3)
    A = torch.randn(n, k, generator=g, dtype=torch.float64)
    C0 = torch.randn(n, n, generator=g, dtype=torch.float64)
    Ad, Cd = _t(A.numpy()), _t(C0.numpy())
    eng.gemm_nt(Ad, Ad, Cd, alpha=-1.0, beta=1.0, tri=True)
    ref = C0 - A @ A.T
    out = Cd.cpu()
    low = torch.tril(torch.ones(n, n, dtype=torch.bool))
    assert (out[low] - ref[low]).abs().max().item() < 1e-10
    assert torch.equal(out[~low], C0[~low])           # strictly upper part bit-identical


# ----------------------------------------------------------------------------------------------- cov build
@pytest.mark.parametrize("name", ["battgp", "rbf_iso", "matern_periodic"])
@pytest.mark.parametrize("n", [1, 63, 64, 129, 1000])
def test_cov_build_symmetric_matches_oracle(eng, name, n):
    es, os_ = _specs()[name]
    x, _ = orc.synth_field_data(n, seed=n)
    if name == "rbf_iso":
        x = np.ascontiguousarray(x[:, 1:4]) / np.array([12.0, 30.0, 40.0])
    K = eng.cov_build(es, _t(x), noise=2.33e-6, symmetric=True).cpu().numpy()
    ref = orc.train_cov(os_, x, 2.33e-6)
    low = np.tril_indices(n)
    # elementwise: |dK| <= 1e-13 * |K| + tiny absolute (exp / sin argument rounding)
    np.testing.assert_allclose(K[low], ref[low], rtol=2e-13, atol=1e-18)


@pytest.mark.parametrize("name", ["battgp", "rbf_iso", "matern_periodic"])
@pytest.mark.parametrize("n1,n2", [(300, 1000), (1, 5), (65, 129), (7, 1)])
def test_cov_build_cross_matches_oracle(eng, name, n1, n2):
    es, os_ = _specs()[name]
    x, _ = orc.synth_field_data(n2, seed=1)
    xq, _ = orc.synth_field_data(n1, seed=2)
    if name == "rbf_iso":
        x, xq = x[:, 1:4] / 20.0, xq[:, 1:4] / 20.0
    K = eng.cov_build(es, _t(xq), _t(x)).cpu().numpy()
    np.testing.assert_allclose(K, orc.cov(os_, xq, x), rtol=2e-13, atol=1e-18)
    d = eng.cov_diag(es, _t(xq)).cpu().numpy()
    np.testing.assert_allclose(d, orc.cov_diag(os_, xq), rtol=1e-14)


def test_cov_build_empty(eng):
    es, _ = _specs()["battgp"]
    x = torch.empty(0, 4, dtype=torch.float64, device=DEV)
    assert eng.cov_build(es, x, noise=1.0, symmetric=True).shape == (0, 0)


def test_cov_build_rejects_bad_spec(eng):
    from battgp_b200 import engine as E, _lib
    bad = E.KernelSpec([E.Term(_lib.RBF, [0], 1.0, (0.0,))])       # zero lengthscale
    with pytest.raises(_lib.BattGPLibraryError):
        eng.cov_build(bad, torch.zeros(4, 1, dtype=torch.float64, device=DEV), symmetric=True)


# ----------------------------------------------------------------------------------------------- potrf & solves
def _spd(n, seed):
    rng = np.random.default_rng(seed)
    x, _ = orc.synth_field_data(n, seed=seed)
    k = orc.train_cov(orc.battgp_spec(), x, 2.33e-6)
    return k, rng


@pytest.mark.parametrize("n,nb", [(1, 1024), (5, 1024), (128, 1024), (129, 1024), (300, 1024), (1000, 1024),
                                  (1300, 512), (2500, 512), (3000, 256)])
def test_potrf_matches_lapack(eng, n, nb):
    k, _ = _spd(n, n)
    ref = sla.cholesky(k, lower=True)
    eng.set("nb", nb)
    try:
        A = _t(k)
        info, logdet, dinv = eng.potrf(A)
    finally:
        eng.set("nb", 0)
    assert info == 0
    L = np.tril(A.cpu().numpy())
    # cond(K) ~ 1e6..1e7 here: backward-stable factor => relative Frobenius error ~ cond * eps
    assert np.linalg.norm(L - ref) / np.linalg.norm(ref) < 1e-9
    assert np.linalg.norm(L @ L.T - k) / np.linalg.norm(k) < 1e-14
    assert abs(logdet - 2 * np.log(np.diag(ref)).sum()) < 1e-9 * max(1.0, abs(logdet))
    # stored inverses of the 128-blocks
    nblk = (n + 127) // 128
    d = dinv.cpu().numpy().reshape(nblk, 128, 128)
    for b in (0, nblk - 1):
        s = b * 128
        e = min(n, s + 128)
        inv_ref = np.linalg.inv(ref[s:e, s:e])
        got = d[b, : e - s, : e - s]
        assert np.linalg.norm(got - inv_ref) / np.linalg.norm(inv_ref) < 1e-9


@pytest.mark.parametrize("M,N,K,tri", [(128, 64, 64, False), (300, 200, 128, False), (1000, 700, 512, False), (1536, 1536, 1024, True)])
def test_gemm_nt_i8_ozaki_is_fp64_accurate(eng, M, N, K, tri):
    """int8-sliced tcgen05 product vs an fp64 reference, with badly scaled rows (per-row exponents)."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp(3 * torch.randn(M, 1, generator=g, dtype=torch.float64))
    B = A if tri else torch.randn(N, K, generator=g, dtype=torch.float64) * torch.exp(3 * torch.randn(N, 1, generator=g, dtype=torch.float64))
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64)
    ref = C0 - A @ B.T
    Ad, Bd, Cd = _t(A.numpy()), (_t(B.numpy()) if not tri else None), _t(C0.numpy())
    if tri:
        Bd = Ad
    eng.gemm_nt_i8(Ad, Bd, Cd, alpha=-1.0, tri=tri)
    out = Cd.cpu()
    scale = A.abs().amax(1, keepdim=True) * B.abs().amax(1, keepdim=True).T * K
    err = (out - ref).abs() / scale
    if tri:
        low = torch.tril(torch.ones(M, N, dtype=torch.bool))
        assert torch.equal(out[~low], C0[~low])
        err = err[low]
    # fp64's own worst-case bound is K * 2^-53 ~ 1e-16 in these units; dropped slice pairs add <= 2^-56 * #pairs
    assert err.max().item() < 4e-16 * 8, err.max().item()


def _oz_model(A, B, C0, alpha):
    """Exact model of csrc/ozaki.cu in Python integers + one rounding per fp64 step of its epilogue: per-row power-of-two
    scale from the row maximum, q = rint(a 2^(55-e)) cut into seven 8-bit two's-complement digits, the 28 digit-plane products
    with s + t <= 6 accumulated per class c = s + t as exact integers, fp64 recombination acc = fma(P_c, 2^-8c, acc),
    tbuf = acc * (alpha 2^eA / 2^14), C = fma(tbuf, 2^eB, C)."""
    import math
    from fractions import Fraction

    def slice_rows(X):
        es, digs = [], []
        for row in X:
            mx = max(abs(float(v)) for v in row)
            e = 0
            if mx > 0:
                hi = np.float64(mx).view(np.uint64) >> np.uint64(32)            # the kernel's upper bound from the high word
                ub = np.uint64((int(hi) << 32) | 0xFFFFFFFF).view(np.float64)
                f, e = math.frexp(float(ub))
                if f > 0.99:
                    e += 1
            q = [int(np.rint(np.ldexp(float(v), 55 - e))) for v in row]
            d = []
            for v in q:                                                        # balanced digits: v = sum_s d_s 256^(6-s), d_s in [-128, 127]
                ds = []
                for _ in range(7):
                    r = ((v + 128) % 256) - 128
                    ds.append(r)
                    v = (v - r) // 256
                d.append(ds[::-1])
            es.append(e); digs.append(d)
        return es, np.array(digs, dtype=object)                                  # [rows, K, 7]

    ea, da = slice_rows(A)
    eb, db = slice_rows(B)
    out = np.array(C0, dtype=np.float64).copy()

    def rnd(fr):                                                                # Fraction -> nearest double (one rounding)
        return float(fr)

    for i in range(A.shape[0]):
        sa = float(alpha) * math.ldexp(1.0, ea[i]) * (1.0 / 16384.0)
        for j in range(B.shape[0]):
            P = [0] * 7
            for s in range(7):
                for t in range(7 - s):
                    P[s + t] += int(sum(int(x) * int(y) for x, y in zip(da[i, :, s], db[j, :, t])))
            acc = float(P[0])
            for c in range(1, 7):
                acc = rnd(Fraction(P[c]) * Fraction(1, 1 << (8 * c)) + Fraction(acc))
            tb = rnd(Fraction(acc) * Fraction(sa))
            out[i, j] = rnd(Fraction(tb) * Fraction(math.ldexp(1.0, eb[j])) + Fraction(float(out[i, j])))
    return out


def test_gemm_nt_i8_is_bit_identical_to_the_integer_model(eng):
    """The digit planes, the int32 class sums and the fp64 recombination are deterministic: the GPU result must equal the
    exact-integer model of _oz_model BIT FOR BIT -- checked on a product small enough for Python integers."""
    rng = np.random.default_rng(3)
    M, N, K = 40, 24, 64
    A = rng.normal(size=(M, K)) * np.exp(2 * rng.normal(size=(M, 1)))
    B = rng.normal(size=(N, K)) * np.exp(2 * rng.normal(size=(N, 1)))
    C0 = rng.normal(size=(M, N))
    Cd = _t(C0)
    eng.gemm_nt_i8(_t(A), _t(B), Cd, alpha=-1.0)
    out = Cd.cpu().numpy()
    # scale bookkeeping of the kernel: a = q 2^(e-55), so a b = q_a q_b 2^(ea+eb-110); class c carries 256^(12-c) = 2^(96-8c);
    # the epilogue applies 2^-14 with the row scale: 2^(96-110) = 2^-14
    ref = _oz_model(A, B, C0, -1.0)
    assert np.array_equal(out, ref), float(np.abs(out - ref).max())


def test_int8_kernel_variants_are_bit_identical(eng):
    """Issue order, relay warp, A-collector, stage fence, L2 hints and the cluster forms only change WHEN things happen; the
    integer products are exact, so every variant must reproduce the default result bit for bit."""
    g = torch.Generator().manual_seed(5)
    n, K = 1536, 512
    A = torch.randn(n, K, generator=g, dtype=torch.float64) * torch.exp(2 * torch.randn(n, 1, generator=g, dtype=torch.float64))
    Ad = _t(A.numpy())
    buf = eng.oz_slice(Ad)
    defaults = {"oz_order": 2, "oz_relay": 1, "oz_collector": 0, "oz_kfence": 1, "oz_l2hint": 0, "oz_cluster": 1}

    def run(**kw):
        for k, v in {**defaults, **kw}.items():
            eng.set(k, v)
        C = torch.zeros(n, n, dtype=torch.float64, device=Ad.device)
        eng.oz_gemm(buf, n, 0, buf, n, 0, C, K, alpha=-1.0, tri=True)
        torch.cuda.synchronize()
        return C

    try:
        base = run()
        ref = -(A @ A.T)
        low = torch.tril(torch.ones(n, n, dtype=torch.bool))
        scale = A.abs().amax(1, keepdim=True) * A.abs().amax(1, keepdim=True).T * K
        assert ((base.cpu() - ref).abs() / scale)[low].max().item() < 4e-16 * 8
        for kw in ({"oz_order": 0}, {"oz_order": 1}, {"oz_relay": 0}, {"oz_collector": 1, "oz_order": 0}, {"oz_kfence": 0},
                   {"oz_l2hint": 1}, {"oz_l2hint": 2}, {"oz_cluster": 2}, {"oz_cluster": 4}):
            assert torch.equal(run(**kw), base), kw
    finally:
        for k, v in defaults.items():
            eng.set(k, v)


def test_gemm_nt_i8_special_rows(eng):
    """zero rows, denormal-sized rows, and NaN / Inf rows (poisoned rows must come out NaN as in fp64 arithmetic)."""
    M, N, K = 256, 128, 128
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    B = torch.randn(N, K, generator=g, dtype=torch.float64)
    A[3] = 0.0
    A[7] *= 1e-290
    B[5] *= 1e+150
    A[11, 17] = float("nan")
    B[9, 3] = float("inf")
    C0 = torch.zeros(M, N, dtype=torch.float64)
    Cd = _t(C0.numpy())
    eng.gemm_nt_i8(_t(A.numpy()), _t(B.numpy()), Cd, alpha=1.0)
    out = Cd.cpu()
    ref = A @ B.T
    assert torch.isnan(out[11]).all() and torch.isnan(out[:, 9]).all()
    ok = torch.ones(M, N, dtype=torch.bool); ok[11] = False; ok[:, 9] = False
    assert torch.equal(out[3][ok[3]], torch.zeros(int(ok[3].sum()), dtype=torch.float64))
    scale = (A.abs().amax(1, keepdim=True) * B.abs().amax(1, keepdim=True).T) * K
    err = ((out - ref).abs() / scale.clamp_min(1e-300))[ok]
    assert torch.isfinite(out[ok]).all() and err.max().item() < 4e-15


@pytest.mark.parametrize("n,nb", [(1300, 256), (2500, 512), (4321, 512), (5000, 1024), (6200, 2048)])
def test_potrf_with_int8_trailing_updates_matches_lapack(eng, n, nb):
    k, _ = _spd(n, n)
    ref = sla.cholesky(k, lower=True)
    res = {}
    for oz in (1, 0):
        eng.set("nb", nb); eng.set("ozaki", oz)
        try:
            A = _t(k)
            info, logdet, _ = eng.potrf(A)
        finally:
            eng.set("nb", 0); eng.set("ozaki", 1)
        assert info == 0
        L = np.tril(A.cpu().numpy())
        res[oz] = (np.linalg.norm(L - ref) / np.linalg.norm(ref), np.linalg.norm(L @ L.T - k) / np.linalg.norm(k),
                   abs(logdet - 2 * np.log(np.diag(ref)).sum()))
    for oz in (1, 0):
        assert res[oz][0] < 1e-9 and res[oz][1] < 1e-14 and res[oz][2] < 1e-9 * abs(2 * np.log(np.diag(ref)).sum()), res
    # the int8 path is as accurate as the DMMA path (same order of magnitude of the forward error)
    assert res[1][0] < 10 * res[0][0] + 1e-13, res


def test_potrf_lookahead_equals_plain_recursion(eng):
    k, _ = _spd(2300, 11)
    A1, A2 = _t(k), _t(k)
    eng.set("nb", 256)
    try:
        i1, ld1, _ = eng.potrf(A1)
        eng.set("lookahead", 0)
        i2, ld2, _ = eng.potrf(A2)
    finally:
        eng.set("lookahead", 1)
        eng.set("nb", 0)
    assert i1 == 0 and i2 == 0
    L1, L2 = torch.tril(A1), torch.tril(A2)
    assert ((L1 - L2).norm() / L2.norm()).item() < 1e-10
    assert abs(ld1 - ld2) < 1e-8 * abs(ld2)


def test_potrf_reports_first_bad_pivot(eng):
    n = 400
    rng = np.random.default_rng(0)
    a = rng.normal(size=(n, n))
    k = a @ a.T + n * np.eye(n)
    k[300, 300] = -1.0                      # leading minor 301 is not PD
    _, info_ref = sla.lapack.dpotrf(k, lower=1)
    info, _, _ = eng.potrf(_t(k))
    assert info == info_ref == 301


@pytest.mark.parametrize("n", [1, 100, 128, 129, 1000, 1500])
def test_potrs_vec_and_lml(eng, n):
    k, rng = _spd(n, 100 + n)
    y = rng.normal(size=n) * 1e-3
    Lr = sla.cholesky(k, lower=True)
    zr = sla.solve_triangular(Lr, y, lower=True)
    ar = sla.solve_triangular(Lr, zr, lower=True, trans="T")
    A = _t(k)
    info, logdet, dinv = eng.potrf(A)
    assert info == 0
    z, alpha = eng.potrs_vec(A, dinv, _t(y))
    assert np.linalg.norm(z.cpu().numpy() - zr) / np.linalg.norm(zr) < 1e-9
    assert np.linalg.norm(alpha.cpu().numpy() - ar) / np.linalg.norm(ar) < 1e-8
    # residual of the normal equations (size-independent property)
    assert np.linalg.norm(k @ alpha.cpu().numpy() - y) / np.linalg.norm(y) < 1e-9
    lml = eng.lml(z, logdet)
    lml_ref = -0.5 * zr @ zr - np.log(np.diag(Lr)).sum() - 0.5 * n * math.log(2 * math.pi)
    assert abs(lml - lml_ref) < 1e-9 * abs(lml_ref)


@pytest.mark.parametrize("n,m", [(1, 1), (128, 3), (129, 300), (1000, 300), (1500, 64)])
def test_trsm_rlt(eng, n, m):
    k, rng = _spd(n, 200 + n)
    w = rng.normal(size=(m, n))
    Lr = sla.cholesky(k, lower=True)
    ref = sla.solve_triangular(Lr, w.T, lower=True).T          # W L^-T
    A = _t(k)
    info, _, dinv = eng.potrf(A)
    assert info == 0
    X = eng.trsm_rlt(A, dinv, _t(w))
    assert np.linalg.norm(X.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-9


# ----------------------------------------------------------------------------------------------- gradient pass
@pytest.mark.parametrize("n", [100, 129, 700, 1500, 4500, 6200])
def test_potri(eng, n):
    k, _ = _spd(n, 300 + n)
    A = _t(k)
    info, _, dinv = eng.potrf(A)
    assert info == 0
    eng.potri(A, dinv)
    kinv = np.tril(A.cpu().numpy())
    kinv = kinv + np.tril(kinv, -1).T
    ref = np.linalg.inv(k)
    assert np.linalg.norm(kinv - ref) / np.linalg.norm(ref) < 3e-7       # cond(K) ~ 1e6-1e7 (n >= 4096: int8 path)
    assert np.linalg.norm(kinv @ k - np.eye(n)) / math.sqrt(n) < 3e-7


@pytest.mark.parametrize("name", ["battgp", "rbf_iso", "matern_periodic"])
def test_lml_grad_matches_oracle(eng, name):
    from battgp_b200 import engine as E
    es, os_ = _specs()[name]
    n = 600
    x, y = orc.synth_field_data(n, seed=4)
    if name == "rbf_iso":
        x = np.ascontiguousarray(x[:, 1:4]) / 20.0
    noise = 2.33e-6 if name != "rbf_iso" else 0.1
    ref = orc.lml_grad(os_, x, y, noise)
    st = E.fit(es, _t(x), _t(y), noise)
    assert abs(st.lml - ref["lml"]) < 1e-9 * abs(ref["lml"])
    eng.potri(st.L, st.dinv)
    g = eng.lml_grad(es, noise, _t(x), st.L, st.alpha).cpu().numpy()
    flat = [ref["noise"]]
    for t in ref["terms"]:
        flat.append(t["outputscale"])
        flat.extend(t["lengthscale"])
        flat.extend(t["period"])
    flat = np.array(flat)
    assert g.shape == flat.shape
    # tr((aa^T - K^-1) dK) suffers cancellation ~ cond(K) * eps relative to |K^-1| |dK|
    scale = np.maximum(np.abs(flat), 1e-6 * np.abs(flat).max())
    assert np.max(np.abs(g - flat) / scale) < 1e-5, (g, flat)
