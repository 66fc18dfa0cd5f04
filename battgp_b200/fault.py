"""Fault evaluation of a battery on the GPU -- the output side of the ``full_gp`` path.

Mirrors /root/reference/src/batt_models/fault_evaluation.py (``get_fault_evaluation``, :20-45) and the array part of
``_calc_fault_probabilities`` (/root/reference/src/batt_models/fault_probabilities.py:37-101): same names, same argument
meaning, same keys / DataFrame columns -- the arithmetic runs in ``bgp_fault_eval`` (csrc/fault.cu).  Inputs may be numpy
arrays (as the reference passes them) or CUDA tensors (e.g. straight from ``CellBatch``); no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib
from . import engine as E


def _eval(r0_cells, r0var_cells, band: float, threshold: float, device=None):
    as_numpy = isinstance(r0_cells, np.ndarray)
    dev = torch.device(device) if device is not None else (torch.device("cuda", torch.cuda.current_device()) if as_numpy
                                                           or not r0_cells.is_cuda else r0_cells.device)
    eng = E.get_engine(dev)
    r0 = torch.as_tensor(np.ascontiguousarray(r0_cells) if as_numpy else r0_cells, dtype=torch.float64).to(dev).contiguous()
    rv = torch.as_tensor(np.ascontiguousarray(r0var_cells) if isinstance(r0var_cells, np.ndarray) else r0var_cells,
                         dtype=torch.float64).to(dev).contiguous()
    if r0.dim() != 2 or r0.shape != rv.shape:
        raise ValueError("r0_cells and r0var_cells must be [n_times, n_cells] arrays of the same shape")
    M, Cn = r0.shape
    if not 2 <= Cn <= 16:
        raise ValueError("2..16 cells supported")
    mc = [torch.empty((M, Cn), dtype=torch.float64, device=dev) for _ in range(5)]
    mv = [torch.empty((M,), dtype=torch.float64, device=dev) for _ in range(2)]
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = eng.L.bgp_fault_eval(eng.h, p(r0), p(rv), M, Cn, Cn, float(band), float(threshold), p(mc[0]), p(mc[1]), p(mc[2]), p(mc[3]),
                              p(mc[4]), p(mv[0]), p(mv[1]), eng._stream())
    _lib.check(rc, "bgp_fault_eval")
    out = {"P_outside_band": mc[0], "P_above_band": mc[1], "P_below_band": mc[2], "r0_mean": mc[3], "P_over_threshold": mc[4],
           "cells_var": mv[0], "weakest_link": mv[1]}
    if as_numpy:
        out = {k: v.cpu().numpy() for k, v in out.items()}
    return out


def get_fault_evaluation(r0_cells, r0var_cells, r0_band_delta: float, r0_upper_threshold: float) -> Dict[str, np.ndarray]:
    """fault_evaluation.py:20-45 (same keys; plus ``r0_mean`` and ``weakest_link``, which the reference derives next)."""
    return _eval(r0_cells, r0var_cells, r0_band_delta, r0_upper_threshold)


def calc_fault_probabilities_from_arrays(t: np.ndarray, r0_cells: np.ndarray, r0var_cells: np.ndarray, cellnumbers: List[int],
                                         cols_mean: List[str], r0_band: float, r0_upper_threshold: float):
    """The DataFrames of fault_probabilities.py ``_calc_fault_probabilities`` (:37-101) from plain arrays: returns
    (r0_fault_df, r0_mean_mean_gp) with the reference's column names and order."""
    import pandas as pd
    ev = _eval(np.asarray(r0_cells, dtype=np.float64), np.asarray(r0var_cells, dtype=np.float64), r0_band, r0_upper_threshold)
    mm = {f"~{col}": ev["r0_mean"][:, i] for i, col in enumerate(cols_mean)}
    mm = pd.DataFrame(mm)
    mm["R0 mean mean_gp"] = mm.mean(axis=1)
    f = {}
    for i, x in enumerate(cellnumbers):
        f[f"R{x} band_i fault prob"] = ev["P_outside_band"][:, i]
        f[f"R_upper{x} band_i fault prob"] = ev["P_above_band"][:, i]
        f[f"R_lower{x} band_i fault prob"] = ev["P_below_band"][:, i]
    for i, x in enumerate(cellnumbers):
        f[f"R{x} thres fault prob"] = ev["P_over_threshold"][:, i]
    f["R0 mean_gp cells var"] = ev["cells_var"]
    f["Weakest_link_stat"] = ev["weakest_link"]
    df = pd.DataFrame(f)
    df.insert(0, "t", np.asarray(t))
    return df, mm
