"""Development probe: one battery = 9 cells x (N, M=300) fit+predict host to host through CellBatch (concurrent + CUDA graph)
against the one-after-another mode and against 9 plain engine.fit/predict calls."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from battgp_b200 import engine as E
from battgp_b200.batch import CellBatch
from battgp_b200.synth import query_grid, synth_field_data

dev = torch.device("cuda", 0)
for n in ([int(a) for a in sys.argv[1:]] or [1000, 2000, 4000]):
    xs, ys, xqs = [], [], []
    for c in range(9):
        x, y = synth_field_data(n, seed=0, cell=c)
        xs.append(x); ys.append(y); xqs.append(query_grid(x))
    spec = E.battgp_spec()
    res = {"op": "cell_batch", "cells": 9, "n": n, "m": 300}
    for mode, kw in (("concurrent_graph", {"concurrent": True}), ("sequential", {"concurrent": False})):
        cb = CellBatch(dev, n_max=n, **kw)
        for _ in range(3):
            out = cb.run(spec, 2.33e-6, xs, ys, xqs)
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            out = cb.run(spec, 2.33e-6, xs, ys, xqs)
        res[mode + "_ms_per_battery"] = round((time.perf_counter() - t0) / reps * 1e3, 3)
        res[mode + "_mean0"] = float(out[0][0, 0])
        if mode == "concurrent_graph":
            res["graph_used"] = cb.graph_replays > 0
        del cb
    xd = [torch.tensor(x, device=dev) for x in xs]; yd = [torch.tensor(y, device=dev) for y in ys]; qd = [torch.tensor(q, device=dev) for q in xqs]
    def plain():
        for c in range(9):
            st = E.fit(spec, xd[c], yd[c], 2.33e-6, xq=qd[c])
            m, v = E.predict(st, qd[c])
            m.cpu(); v.cpu()
    for _ in range(3):
        plain()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        plain()
    res["nine_plain_fit_predict_calls_ms"] = round((time.perf_counter() - t0) / 10 * 1e3, 3)
    print(json.dumps(res), flush=True)
