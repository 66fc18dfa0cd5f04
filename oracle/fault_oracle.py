"""CPU oracle of the fault evaluation -- TEST INFRASTRUCTURE, NOT PRODUCT (only tests/ import it).

numpy restatement of /root/reference/src/batt_models/fault_evaluation.py:
  normal_cdf .............................. :7-8     0.5 + 0.5 erf((x - mean) / (sqrt 2 std))
  hodges_lehmann_estimator ................ :11-17   median of the Walsh averages (x_i + x_j)/2, i <= j
  calc_outside_band_probabilities ......... :48-91   leave-one-out location, band probabilities
  calc_over_threshold_probability ......... :94-101
  calc_r0_cells_var ....................... :103-104
and of the weakest-link statistic of fault_probabilities.py:88-95.
PINNED: tests/golden/fault_vectors.npz holds the outputs of the reference's own functions (imported, not copied, by
tests/golden/make_fault_golden.py) on seeded inputs; tests/test_fault.py checks this file against them."""
import math

import numpy as np


def normal_cdf(x, mean, std):
    erf = np.vectorize(math.erf)
    return 0.5 + 0.5 * erf((x - mean) / (np.sqrt(2) * std))


def hodges_lehmann(x):
    n = len(x)
    return np.median([(x[i] + x[j]) / 2 for i in range(n) for j in range(i, n)])


def fault_evaluation(r0, r0var, band, thr):
    r0 = np.asarray(r0, np.float64)
    r0var = np.asarray(r0var, np.float64)
    M, C = r0.shape
    r0_mean = np.zeros_like(r0)
    for c in range(C):
        mask = np.ones(C, bool)
        mask[c] = False
        for i in range(M):
            r0_mean[i, c] = hodges_lehmann(r0[i, mask])
    std = np.sqrt(r0var)
    above = 1 - normal_cdf(r0_mean + band, r0, std)
    below = normal_cdf(r0_mean - band, r0, std)
    return {"P_outside_band": above + below, "P_above_band": above, "P_below_band": below, "r0_mean": r0_mean,
            "P_over_threshold": 1 - normal_cdf(thr, r0, std), "cells_var": r0.var(axis=1),
            "weakest_link": 1 - np.prod(1 - (above + below), axis=1)}
