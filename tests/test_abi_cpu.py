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
