"""EXPERIMENTAL end-to-end modular (CRT) int8 product (csrc/next/ozaki2_mma.cu: residue image -> tcgen05 kind::i8 products, two
moduli per TMEM pass -> reconstruction) against the CPU model, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import ozaki2_model as oz2                                             # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("m,n,k,spread", [(128, 256, 64, 0.0), (200, 300, 512, 3.0), (384, 512, 2048, 10.0), (5, 7, 128, 0.0)])
def test_modular_gemm_matches_model_bit_for_bit(eng, m, n, k, spread):
    rng = np.random.default_rng(11)
    a = rng.standard_normal((m, k)) * np.exp(rng.uniform(-spread, spread, (m, 1)))
    b = rng.standard_normal((n, k)) * np.exp(rng.uniform(-spread, spread, (n, 1)))
    a[0, :] = np.abs(a[0, :]); b[0, :] = np.abs(b[0, :])
    c0 = rng.standard_normal((m, n))
    c = torch.tensor(c0, device=DEV)
    eng.oz2_gemm(torch.tensor(a, device=DEV), torch.tensor(b, device=DEV), c, alpha=-1.0)
    torch.cuda.synchronize()
    ref = c0 - oz2.gemm_nt_modular(a, b)
    got = c.cpu().numpy()
    assert np.array_equal(got, ref), float(np.max(np.abs(got - ref)))
    plain = c0 - a @ b.T
    scale = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] * k
    assert np.max(np.abs(got - plain) / scale) < 1e-14
