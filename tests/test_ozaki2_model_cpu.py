"""CPU: the modular (CRT) int8 emulation model in tools/ozaki2_model.py -- design groundwork for the next int8 kernel
(DESIGN.md section 5).  Pins the constants and shows the reconstruction is exact and the result fp64-accurate."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import ozaki2_model as oz2                                             # noqa: E402


def test_constants_cover_k_2048():
    oz2.check_constants(2048)
    assert len(oz2.MODULI) == 16 and max(oz2.MODULI) <= 256
    assert 124 < np.log2(float(oz2.P)) < 126


def test_reconstruction_is_exact_against_python_integers():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((6, 2048)) * np.exp(rng.uniform(-20, 20, (6, 1)))
    b = rng.standard_normal((5, 2048)) * np.exp(rng.uniform(-20, 20, (5, 1)))
    a[0, :] = np.abs(a[0, :])                      # worst case for |C'|: all terms of one sign
    b[0, :] = np.abs(b[0, :])
    c_exact, ea, eb = oz2.exact_scaled_product(a, b)
    g = oz2.int8_products(oz2.residues(oz2.to_scaled_int(a, ea)), oz2.residues(oz2.to_scaled_int(b, eb)))
    cp = oz2.crt_reconstruct(g)
    for i in range(6):
        for j in range(5):
            ref = float(c_exact[i][j])             # correctly rounded
            assert abs(cp[i, j] - ref) <= 2.0 ** -51 * abs(ref) + 0.0, (i, j, cp[i, j], ref)


def test_product_is_fp64_accurate_with_cancellation():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((40, 1024))
    b = rng.standard_normal((33, 1024))
    b[3] = a[7] * 1e-3 + 1e-12 * rng.standard_normal(1024)            # ordinary entries
    c = oz2.gemm_nt_modular(a, b)
    ref = np.array([[float(np.sum(np.array(a[i], dtype=np.longdouble) * np.array(b[j], dtype=np.longdouble)))
                     for j in range(33)] for i in range(40)])
    scale = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] * 1024
    assert np.max(np.abs(c - ref) / scale) < 2.0 ** -52          # truncation: 2 * 2^-54 relative to row maxima, per term
    # and no worse than the plain fp64 product on ordinary data
    assert np.max(np.abs(c - ref)) <= 4 * np.max(np.abs(a @ b.T - ref)) + 1e-300


def test_zero_rows_and_tiny_values():
    a = np.zeros((3, 64)); b = np.zeros((2, 64))
    a[1, 5] = 3.0; b[0, 5] = -2.5e-300; b[1, 7] = 1.0
    c = oz2.gemm_nt_modular(a, b)
    assert c[1, 0] == 3.0 * -2.5e-300 and np.all(c[0] == 0) and c[1, 1] == 0.0
