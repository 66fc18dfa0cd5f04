"""Fault evaluation: the numpy oracle against golden vectors produced by the reference's own
fault_evaluation.py / fault_probabilities.py (CPU), and the CUDA kernel bgp_fault_eval against both (GPU)."""
import os

import numpy as np
import pytest

from oracle import fault_oracle as fo

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fault_vectors.npz"))
KEYS = ["P_outside_band", "P_above_band", "P_below_band", "r0_mean", "P_over_threshold", "cells_var", "weakest_link"]


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_matches_the_reference_outputs(name):
    ev = fo.fault_evaluation(G[f"{name}_r0"], G[f"{name}_r0var"], float(G[f"{name}_band"]), float(G[f"{name}_thr"]))
    for k in KEYS:
        np.testing.assert_allclose(ev[k], G[f"{name}_{k}"], rtol=1e-13, atol=1e-15, err_msg=k)
    # the leave-one-out location is pure add / halve / median: bit-exact
    assert np.array_equal(ev["r0_mean"], G[f"{name}_r0_mean"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_gpu_fault_eval_matches_reference_outputs(eng, name):
    import torch
    from battgp_b200 import fault
    r0, rv = G[f"{name}_r0"], G[f"{name}_r0var"]
    band, thr = float(G[f"{name}_band"]), float(G[f"{name}_thr"])
    ev = fault.get_fault_evaluation(r0, rv, band, thr)                       # numpy in -> numpy out, reference keys
    for k in ("P_outside_band", "P_above_band", "P_below_band", "P_over_threshold", "cells_var"):
        assert isinstance(ev[k], np.ndarray)
    for k in KEYS:
        np.testing.assert_allclose(ev[k], G[f"{name}_{k}"], rtol=1e-12, atol=2e-15, err_msg=k)
    assert np.array_equal(ev["r0_mean"], G[f"{name}_r0_mean"])               # bit-exact integer-like work
    # device tensors in -> device tensors out (results of CellBatch never leave the GPU)
    evd = fault.get_fault_evaluation(torch.tensor(r0, device="cuda:0"), torch.tensor(rv, device="cuda:0"), band, thr)
    assert evd["weakest_link"].is_cuda
    np.testing.assert_allclose(evd["weakest_link"].cpu().numpy(), G[f"{name}_weakest_link"], rtol=1e-12, atol=2e-15)
    # the reference's DataFrames (fault_probabilities.py:37-101): same columns, same order, same numbers
    C = r0.shape[1]
    cells = list(range(1, C + 1))
    df, mm = fault.calc_fault_probabilities_from_arrays(G[f"{name}_t"], r0, rv, cells, [f"r0_acausal_c{c}" for c in cells], band, thr)
    assert list(df.columns) == [str(c) for c in G[f"{name}_df_columns"]]
    assert list(mm.columns) == [str(c) for c in G[f"{name}_mm_columns"]]
    np.testing.assert_allclose(df.to_numpy(dtype=np.float64), G[f"{name}_df_values"], rtol=1e-12, atol=2e-15)
    np.testing.assert_allclose(mm.to_numpy(dtype=np.float64), G[f"{name}_mm_values"], rtol=1e-12, atol=2e-15)


@pytest.mark.gpu
def test_gpu_fault_eval_rejects_bad_shapes(eng):
    from battgp_b200 import fault
    with pytest.raises(ValueError):
        fault.get_fault_evaluation(np.zeros((4, 1)), np.ones((4, 1)), 0.1, 1.0)
    with pytest.raises(ValueError):
        fault.get_fault_evaluation(np.zeros((4, 3)), np.ones((5, 3)), 0.1, 1.0)
