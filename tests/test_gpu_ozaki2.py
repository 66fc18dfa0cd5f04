"""Components of the modular (CRT) int8 emulation (csrc/ozaki2.cu) against the CPU model tools/ozaki2_model.py: the residue
kernel must reproduce the model's int8 residues and exponents exactly, and the reconstruction kernel must return the model's
result bit for bit when fed the exact int32 products (formed here by an exact fp64 torch matmul of the residues)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import ozaki2_model as oz2                                             # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(seed, m, n, k, spread):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((m, k)) * np.exp(rng.uniform(-spread, spread, (m, 1)))
    b = rng.standard_normal((n, k)) * np.exp(rng.uniform(-spread, spread, (n, 1)))
    a[0, :] = np.abs(a[0, :]); b[0, :] = np.abs(b[0, :])      # largest possible |C'|
    if m > 2:
        a[2, :] = 0.0                                          # an all-zero row
    return a, b


@pytest.mark.parametrize("m,n,k,spread", [(5, 7, 64, 0.0), (130, 70, 512, 5.0), (64, 257, 2048, 30.0)])
def test_residues_match_model(eng, m, n, k, spread):
    a, _ = _case(3, m, n, k, spread)
    res, expo = eng.oz2_residues(torch.tensor(a, device=DEV))
    e_ref = oz2.row_scale_exponent(a)
    r_ref = oz2.residues(oz2.to_scaled_int(a, e_ref))
    assert np.array_equal(expo.cpu().numpy().astype(np.int64), e_ref)
    assert np.array_equal(res.cpu().numpy(), r_ref)


@pytest.mark.parametrize("m,n,k,spread", [(5, 7, 64, 0.0), (130, 70, 512, 5.0), (64, 257, 2048, 30.0)])
def test_crt_reconstruction_matches_model_bit_for_bit(eng, m, n, k, spread):
    a, b = _case(4, m, n, k, spread)
    ra, ea = eng.oz2_residues(torch.tensor(a, device=DEV))
    rb, eb = eng.oz2_residues(torch.tensor(b, device=DEV))
    # exact int32 products: |sum| <= K * 2^14 < 2^53, so an fp64 matmul of the residues is exact
    g = torch.matmul(ra.double(), rb.double().transpose(1, 2)).round().to(torch.int32).contiguous()
    assert np.array_equal(g.cpu().numpy(), oz2.int8_products(ra.cpu().numpy(), rb.cpu().numpy()))
    c0 = np.random.default_rng(5).standard_normal((m, n))
    c = torch.tensor(c0, device=DEV)
    eng.oz2_crt(g, ea, eb, c, alpha=-1.0)
    ref = c0 - oz2.gemm_nt_modular(a, b)
    assert np.array_equal(c.cpu().numpy(), ref)
    # and the emulated product is fp64-accurate
    exact = np.array([[float(np.sum(a[i].astype(np.longdouble) * b[j].astype(np.longdouble))) for j in range(n)] for i in range(m)]) \
        if m * n <= 1000 else None
    if exact is not None:
        scale = np.abs(a).max(axis=1)[:, None] * np.abs(b).max(axis=1)[None, :] * k
        assert np.max(np.abs((c0 - c.cpu().numpy()) - exact) / np.maximum(scale, 1e-300)) < 2.0 ** -52


def test_nan_row_poisons_only_its_row(eng):
    a, b = _case(6, 9, 8, 64, 0.0)
    a[4, 10] = np.nan
    ra, ea = eng.oz2_residues(torch.tensor(a, device=DEV))
    rb, eb = eng.oz2_residues(torch.tensor(b, device=DEV))
    g = torch.matmul(ra.double(), rb.double().transpose(1, 2)).round().to(torch.int32).contiguous()
    c = torch.zeros((9, 8), dtype=torch.float64, device=DEV)
    eng.oz2_crt(g, ea, eb, c, alpha=1.0)
    out = c.cpu().numpy()
    assert np.all(np.isnan(out[4])) and np.all(np.isfinite(np.delete(out, 4, axis=0)))
