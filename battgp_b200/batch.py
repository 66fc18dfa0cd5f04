"""Per-cell batch on one GPU: the 9 exact GPs of a battery (pack + 8 cells) share the query grid, the hyper-parameters and
most of their input columns (/root/reference/src/batt_models/battgp_full.py:41-60 builds them, :100-120 predicts each at
300 time points and frees it).  This helper stages ALL cells' inputs with one pinned host buffer / one H2D copy, re-uses one
K buffer and the int8 workspace across the cells (the reference re-allocates per model), runs every cell as a fused
fit+predict (bgp_potrf_aug) and returns all results with a single D2H copy.  Across GPUs the mapping stays the reference's:
one process per GPU, cells round-robin (bench.py --gpus N)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import engine as E


class CellBatch:
    def __init__(self, device, n_max: int, m_query: int = 300):
        self.device = torch.device(device)
        self.eng = E.get_engine(self.device)
        self.n_max, self.m = int(n_max), int(m_query)
        self.K = E.alloc_matrix(self.n_max + self.m, self.n_max, self.device)     # reused by every cell

    def _kview(self, n: int) -> torch.Tensor:
        """(n + m) x n view of the shared buffer with an even leading dimension."""
        ld = n + (n & 1)
        return torch.as_strided(self.K, (n + self.m, n), (ld, 1))

    def run(self, spec: E.KernelSpec, noise: float, xs: Sequence[np.ndarray], ys: Sequence[np.ndarray],
            xqs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, List[float]]:
        """xs[c]: [n_c, D], ys[c]: [n_c], xqs[c]: [m, D].  Returns (means [C, m], variances [C, m], lml per cell)."""
        C = len(xs)
        if not (len(ys) == len(xqs) == C):
            raise ValueError("xs, ys, xqs must have the same length")
        D = xs[0].shape[1]
        sizes = [int(x.shape[0]) for x in xs]
        if max(sizes) > self.n_max or any(xq.shape[0] != self.m for xq in xqs):
            raise ValueError("cell larger than n_max or wrong query count")
        # one pinned staging buffer: [x_0 | y_0 | xq_0 | x_1 | ...]
        total = sum(n * D + n + self.m * D for n in sizes)
        host = torch.empty(total, dtype=torch.float64).pin_memory()
        off, slots = 0, []
        for x, y, xq, n in zip(xs, ys, xqs, sizes):
            a = off; host[a:a + n * D] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(-1)); off += n * D
            b = off; host[b:b + n] = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64).reshape(-1)); off += n
            c = off; host[c:c + self.m * D] = torch.from_numpy(np.ascontiguousarray(xq, dtype=np.float64).reshape(-1)); off += self.m * D
            slots.append((a, b, c, n))
        dev = host.to(self.device, non_blocking=True)
        out = torch.empty((2, C, self.m), dtype=torch.float64, device=self.device)
        lmls = []
        for ci, (a, b, c, n) in enumerate(slots):
            x = dev[a:a + n * D].view(n, D)
            y = dev[b:b + n]
            xq = dev[c:c + self.m * D].view(self.m, D)
            st = E.fit(spec, x, y, noise, K_out=self._kview(n), xq=xq)
            mean, var = E.predict(st, xq)
            out[0, ci].copy_(mean)
            out[1, ci].copy_(var)
            lmls.append(st.lml)
        res = out.cpu().numpy()
        return res[0], res[1], lmls
