"""Generates tests/golden/fault_vectors.npz by IMPORTING the reference's own fault code (build container only):
/root/reference/src/batt_models/fault_evaluation.py (pure numpy/scipy) and fault_probabilities.py::_calc_fault_probabilities
(through the import shim, because its package imports gpytorch).  Seeded inputs shaped like a battery's result: M = 300 query
times x 8 cells, r0 around 1.1 mOhm with one drifting cell, GP variances 2e-8 .. 1e-7."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import battgp_b200.shim as shim  # noqa: E402

shim.install(force=True)
sys.path.insert(0, "/root/reference")
import pandas as pd  # noqa: E402
from src.batt_models import fault_evaluation as fe  # noqa: E402
from src.batt_models.fault_probabilities import _calc_fault_probabilities  # noqa: E402

rng = np.random.default_rng(42)
out = {}
for name, (M, C) in {"a": (300, 8), "b": (37, 5), "c": (64, 2)}.items():
    t = np.linspace(0, 117, M)
    r0 = 1.1e-3 + 2e-5 * rng.normal(size=(M, C)) + 1e-6 * t[:, None]
    r0[:, C - 1] += 3e-6 * t                                   # one cell drifts away
    r0var = rng.uniform(2e-8, 1e-7, size=(M, C)) * 1e-3        # (sigma ~ 5e-6 .. 1e-5 Ohm)
    band, thr = 0.25e-3 * 0.2, 1.3e-3
    ev = fe.get_fault_evaluation(r0, r0var, band, thr)
    pb = fe.calc_outside_band_probabilities(r0, r0var, band, band_mean_without_eval_cell=True, return_intermediate_arrays=True)
    cells = list(range(1, C + 1))
    dfm = pd.DataFrame({"t": t, **{f"r0_acausal_c{c}": r0[:, i] for i, c in enumerate(cells)}})
    dfv = pd.DataFrame({"t": t, **{f"r0var_acausal_c{c}": r0var[:, i] for i, c in enumerate(cells)}})
    df_f, df_mm = _calc_fault_probabilities(dfm, dfv, cells, band, thr)
    out.update({f"{name}_t": t, f"{name}_r0": r0, f"{name}_r0var": r0var, f"{name}_band": band, f"{name}_thr": thr,
                f"{name}_P_outside_band": ev["P_outside_band"], f"{name}_P_above_band": ev["P_above_band"],
                f"{name}_P_below_band": ev["P_below_band"], f"{name}_P_over_threshold": ev["P_over_threshold"],
                f"{name}_cells_var": ev["cells_var"], f"{name}_r0_mean": pb[3],
                f"{name}_weakest_link": df_f["Weakest_link_stat"].to_numpy(),
                f"{name}_df_columns": np.array(list(df_f.columns)), f"{name}_df_values": df_f.to_numpy(dtype=np.float64),
                f"{name}_mm_columns": np.array(list(df_mm.columns)), f"{name}_mm_values": df_mm.to_numpy(dtype=np.float64)})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fault_vectors.npz"), **out)
print("written", len(out), "arrays")
