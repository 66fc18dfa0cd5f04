"""CPU-side checks (no GPU): the C-ABI library loads, exports every symbol include/battgp_b200.h declares, the ctypes
mirror of the structs matches the header, and the product path refuses to run without CUDA (no silent fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from battgp_b200 import _lib, engine as E


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), name
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert L.bgp_version() == 100


def test_struct_layout_matches_header():
    hdr = open(_lib.HEADER).read()
    assert int(re.search(r"#define BGP_MAX_TERMS\s+(\d+)", hdr).group(1)) == _lib.MAX_TERMS
    assert int(re.search(r"#define BGP_MAX_DIMS\s+(\d+)", hdr).group(1)) == _lib.MAX_DIMS
    # int32 type, ndims, dims[8]; double outputscale, lengthscale[8], period[8]
    assert C.sizeof(_lib.BgpTerm) == 4 + 4 + 32 + 8 + 64 + 64
    assert C.sizeof(_lib.BgpKernelSpec) == 8 + 4 * C.sizeof(_lib.BgpTerm) + 8


def test_pure_host_entry_points():
    L = _lib.lib()
    assert L.bgp_potrf_dinv_elems(0) == 0
    assert L.bgp_potrf_dinv_elems(1) == 128 * 128
    assert L.bgp_potrf_dinv_elems(129) == 2 * 128 * 128
    cs = E.battgp_spec().to_c(2.33e-6)
    assert L.bgp_grad_slots(C.byref(cs)) == 1 + 1 + (1 + 3)
    cs = E.matern_periodic_spec().to_c(0.0)
    assert L.bgp_grad_slots(C.byref(cs)) == 1 + (1 + 3) + (1 + 1 + 1)
    bad = _lib.BgpKernelSpec()
    bad.nterms = 9
    assert L.bgp_grad_slots(C.byref(bad)) == _lib.E_SPEC


def test_kernel_spec_marshalling():
    cs = E.battgp_spec(1e-12, 0.01, (1.0, 2.0, 3.0)).to_c(0.5)
    assert cs.nterms == 2 and cs.noise == 0.5
    assert cs.terms[0].type == _lib.WIENER and list(cs.terms[0].dims)[:1] == [0]
    assert cs.terms[1].type == _lib.RBF and list(cs.terms[1].dims)[:3] == [1, 2, 3]
    assert list(cs.terms[1].lengthscale)[:3] == [1.0, 2.0, 3.0]
    iso = E.KernelSpec([E.Term(_lib.RBF, [0, 1, 2], 3.0, (2.0,))]).to_c()
    assert list(iso.terms[0].lengthscale)[:3] == [2.0, 2.0, 2.0]
    with pytest.raises(ValueError):
        E.KernelSpec([]).to_c()
    with pytest.raises(ValueError):
        E.KernelSpec([E.Term(_lib.RBF, [0, 1], 1.0, (1.0, 2.0, 3.0))]).to_c()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    with pytest.raises(_lib.BattGPLibraryError):
        E.get_engine(torch.device("cpu"))
    x = torch.zeros(4, 4, dtype=torch.float64)
    with pytest.raises(_lib.BattGPLibraryError):
        E.fit(E.battgp_spec(), x, torch.zeros(4, dtype=torch.float64), 1e-3)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIBPATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.BattGPLibraryError):
        _lib.lib()


def test_product_code_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "battgp_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def _schedule(rows, n, ozaki=1, nb=0):
    L = _lib.lib()
    buf = (C.c_int64 * 4096)()
    np_ = L.bgp_panel_schedule(rows, n, ozaki, nb, buf, 4096)
    assert np_ >= 0
    return [int(buf[i]) for i in range(np_ + 1)]


def test_panel_schedule_covers_the_columns_and_narrows_towards_the_end():
    """Host logic of the look-ahead Cholesky (api.cu make_schedule; no CUDA call): panels tile [0, n) exactly, every panel but
    the last is a multiple of 128 wide, widths follow the remaining rows (2048 / 1024 / 512 on the int8 path, never above 1024
    on the DMMA-only path) and never grow again, and the "nb" knob gives uniform panels."""
    for rows, n in ((40300, 40000), (40000, 40000), (16384, 16384), (8192, 8192), (3300, 3000), (1000, 1000), (129, 129), (5, 5)):
        for oz in (0, 1):
            s = _schedule(rows, n, oz)
            assert s[0] == 0 and s[-1] == n and all(b > a for a, b in zip(s, s[1:]))
            w = [b - a for a, b in zip(s, s[1:])]
            assert all(x % 128 == 0 for x in w[:-1])
            assert all(w[i + 1] <= w[i] for i in range(len(w) - 1))
            assert max(w) <= (2048 if oz else 1024)
    w = [b - a for a, b in zip(*(lambda s: (s, s[1:]))(_schedule(40300, 40000, 1)))]
    assert w[0] == 2048 and 1024 in w and w[-2] == 512
    assert _schedule(8192, 8192, 1) == list(range(0, 8193, 512))
    s = _schedule(5000, 5000, 1, nb=1024)
    assert s == [0, 1024, 2048, 3072, 4096, 5000]
    assert _schedule(0, 0) == [0]
