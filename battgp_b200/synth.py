"""Synthetic 8s1p field telemetry for benchmarks: statistics mirrored from the reference's own
fixtures (tests/data/cache/{3,14}.feather after the reference's filter) -- t in [0,120] d on a 5 s grid, I in
[-80,-5] A, SOC in [40,95] %, T ~ N(24,4) clipped to [10,35] C, R ~ 4 mOhm with a slow ageing trend and
sigma_n^2 = 2.33e-6 (config.py:39).  X is [N,4] = [t_days, I, SOC, T] as produced by
/root/reference/src/batt_data/batt_data.py:248-256."""
from __future__ import annotations

import math

import numpy as np


def synth_field_data(n: int, seed: int = 0, cell: int = 0):
    rng = np.random.default_rng(seed)
    t = np.sort(np.round(rng.uniform(0, 120, n) * 17280) / 17280)
    cur = rng.uniform(-80, -5, n)
    soc = rng.uniform(40, 95, n)
    temp = np.clip(rng.normal(24, 4, n), 10, 35)
    noise = rng.normal(0, math.sqrt(2.33e-6), n)
    if cell:  # config 4: per-cell perturbation of I, T and an offset on y (same t, SOC)
        rc = np.random.default_rng(1000 + cell)
        cur = cur + rc.uniform(-2, 2, n)
        temp = temp + rc.normal(0, 0.5, n)
    x = np.ascontiguousarray(np.stack([t, cur, soc, temp], axis=1))
    y = (4e-3 * (1 + 2e-3 * t) + 1e-3 * np.exp(-(temp - 10) / 15) + 5e-4 * (soc - 70) ** 2 / 900
         - 1e-5 * cur / 80 + noise + cell * 1e-4)
    return x, y


def query_grid(x: np.ndarray, m: int = 300, op=(-15.0, 90.0, 25.0)) -> np.ndarray:
    """battgp_full.py:98 + battcellgp_full.py:199-206; op = gp_runner.py:32."""
    t = np.linspace(x[0, 0], x[-1, 0], m)
    return np.column_stack([t, np.full(m, op[0]), np.full(m, op[1]), np.full(m, op[2])])
