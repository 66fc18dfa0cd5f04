"""Per-cell batch on one GPU: the 9 exact GPs of a battery (pack + 8 cells) share the query grid, the hyper-parameters and
most of their input columns (/root/reference/src/batt_models/battgp_full.py:41-60 builds them, :100-120 predicts each at
300 time points and frees it).  Across GPUs the mapping stays the reference's: one process per GPU (bench.py --gpus N).

Two modes:

* **large cells** (n_max > SMALL_N): the cells run one after another -- each one fills the GPU -- with ONE pinned host buffer /
  one H2D copy for all inputs, one shared K buffer and int8 workspace (the reference re-allocates per model), the fused
  fit+predict per cell (bgp_potrf_aug) and one D2H copy.

* **small cells** (n_max <= SMALL_N, the reference's default N = 1000, config.py:29): one such GP is latency-bound (a chain of
  ~100 small kernels on a few SMs), so the cells run CONCURRENTLY -- one stream and one engine context per cell, fully
  asynchronous entry points (bgp_potrf_async / bgp_lml_dev: no host read-back inside the chain) -- and the whole battery
  (H2D copy, 9 x [build, factorise, solve, LML, predict], D2H copy) is captured ONCE in a CUDA graph and replayed for every
  further battery with the same cell sizes and hyper-parameters: a battery costs one graph launch instead of ~700 kernel
  launches.  A cell whose factorisation reports a non-positive pivot is redone through engine.fit (GPyTorch's jitter retries).

Shared sub-terms: the 8 cells share t and SOC, so the Wiener term and the SOC factor of the RBF term are identical across
cells.  They are deliberately NOT computed once and re-read: the build is bound by FP64 issue / HBM writes (8 B per entry out),
a shared N x N term would add 8-16 B of reads per entry to save ~15 of ~45 flops -- it costs more HBM time than it saves
(DESIGN.md section 6).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import engine as E

SMALL_N = 4096
INT_MAX = 2 ** 31 - 1


class CellBatch:
    def __init__(self, device, n_max: int, m_query: int = 300, concurrent: bool | None = None, max_streams: int = 9):
        self.device = torch.device(device)
        self.eng = E.get_engine(self.device)
        self.n_max, self.m = int(n_max), int(m_query)
        self.concurrent = (self.n_max <= SMALL_N) if concurrent is None else bool(concurrent)
        self.max_streams = int(max_streams)
        self.K = None if self.concurrent else E.alloc_matrix(self.n_max + self.m, self.n_max, self.device)   # reused by every cell
        self._slots = []          # concurrent mode: (Engine, stream, K buffer)
        self._graphs = {}         # key -> captured battery
        self.graph_replays = 0
        self.use_graph = True

    def _kview(self, K: torch.Tensor, n: int) -> torch.Tensor:
        """(n + m) x n view of a K buffer with an even leading dimension."""
        ld = n + (n & 1)
        return torch.as_strided(K, (n + self.m, n), (ld, 1))

    # ------------------------------------------------------------------------------------------------ staging
    def _layout(self, sizes: Sequence[int], D: int):
        off, slots = 0, []
        for n in sizes:
            a = off; off += n * D
            b = off; off += n
            c = off; off += self.m * D
            slots.append((a, b, c, n))
        return off, slots

    def _fill(self, host: torch.Tensor, slots, xs, ys, xqs, D: int):
        for (a, b, c, n), x, y, xq in zip(slots, xs, ys, xqs):
            host[a:a + n * D] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(-1))
            host[b:b + n] = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64).reshape(-1))
            host[c:c + self.m * D] = torch.from_numpy(np.ascontiguousarray(xq, dtype=np.float64).reshape(-1))

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
        if self.concurrent:
            return self._run_concurrent(spec, noise, xs, ys, xqs, sizes, D)
        # one pinned staging buffer: [x_0 | y_0 | xq_0 | x_1 | ...]
        total, slots = self._layout(sizes, D)
        host = torch.empty(total, dtype=torch.float64).pin_memory()
        self._fill(host, slots, xs, ys, xqs, D)
        dev = host.to(self.device, non_blocking=True)
        out = torch.empty((2, C, self.m), dtype=torch.float64, device=self.device)
        lmls = []
        for ci, (a, b, c, n) in enumerate(slots):
            x = dev[a:a + n * D].view(n, D)
            y = dev[b:b + n]
            xq = dev[c:c + self.m * D].view(self.m, D)
            st = E.fit(spec, x, y, noise, K_out=self._kview(self.K, n), xq=xq)
            mean, var = E.predict(st, xq)
            out[0, ci].copy_(mean)
            out[1, ci].copy_(var)
            lmls.append(st.lml)
        res = out.cpu().numpy()
        return res[0], res[1], lmls

    # ------------------------------------------------------------------------------------------------ small cells
    def _ensure_slots(self, count: int):
        while len(self._slots) < min(count, self.max_streams):
            eng = E.Engine(self.device)               # own bgp_ctx: own scratch scalars, safe to run beside the others
            eng.set("pdl", 0)                         # plain stream order inside the captured graph
            # concurrent cells fill the GPU, so throughput counts, not latency: above ~1.5k rows the panel schedule (wide-K
            # int8 updates) does less work per cell than one leaf chain over the whole matrix (9 x N=4000: 13.4 vs 14.7 ms)
            eng.set("chain_whole_max", 1536)
            self._slots.append((eng, torch.cuda.Stream(self.device), E.alloc_matrix(self.n_max + self.m, self.n_max, self.device)))

    def _capture(self, spec, noise, sizes, D):
        C = len(sizes)
        total, slots = self._layout(sizes, D)
        self._ensure_slots(C)
        g = {"host_in": torch.empty(total, dtype=torch.float64).pin_memory(),
             "dev_in": torch.empty(total, dtype=torch.float64, device=self.device),
             "out": torch.empty((2 * C * self.m + 2 * C,), dtype=torch.float64, device=self.device),
             "info": torch.empty((C,), dtype=torch.int32, device=self.device),
             "host_out": torch.empty((2 * C * self.m + 2 * C,), dtype=torch.float64).pin_memory(),
             "host_info": torch.empty((C,), dtype=torch.int32).pin_memory(), "slots": slots}
        m = self.m
        means = g["out"][:C * m].view(C, m)
        vars_ = g["out"][C * m:2 * C * m].view(C, m)
        lml = g["out"][2 * C * m:2 * C * m + C]
        logdet = g["out"][2 * C * m + C:]

        def enqueue():
            main = torch.cuda.current_stream(self.device)
            g["dev_in"].copy_(g["host_in"], non_blocking=True)
            fork = torch.cuda.Event()
            fork.record(main)
            used = []
            for ci, (a, b, c, n) in enumerate(slots):
                eng, st, K = self._slots[ci % len(self._slots)]
                if st not in used:
                    st.wait_event(fork)
                    used.append(st)
                with torch.cuda.stream(st):
                    x = g["dev_in"][a:a + n * D].view(n, D)
                    y = g["dev_in"][b:b + n]
                    xq = g["dev_in"][c:c + m * D].view(m, D)
                    Kf = self._kview(K, n)
                    eng.cov_build(spec, x, noise=noise, symmetric=True, out=Kf[:n])
                    eng.cov_build(spec, xq, x, out=Kf[n:])
                    dinv = eng.potrf_async(Kf, g["info"][ci:ci + 1], logdet[ci:ci + 1])     # query rows leave as V = K_*N L^-T
                    z, alpha = eng.potrs_vec(Kf[:n], dinv, y)
                    eng.lml_dev(z, logdet[ci:ci + 1], lml[ci:ci + 1])
                    Kq = eng.cov_build(spec, xq, x)
                    eng.predict_tail(Kq=Kq, alpha=alpha, mean_out=means[ci])
                    eng.predict_tail(V=Kf[n:], kdiag=eng.cov_diag(spec, xq), var_out=vars_[ci])
            for st in used:
                join = torch.cuda.Event()
                join.record(st)
                main.wait_event(join)
            g["host_out"].copy_(g["out"], non_blocking=True)
            g["host_info"].copy_(g["info"], non_blocking=True)

        g["enqueue"] = enqueue
        return g

    def _run_concurrent(self, spec, noise, xs, ys, xqs, sizes, D):
        C, m = len(sizes), self.m
        cs = spec.to_c(noise)
        key = (tuple(sizes), D, bytes(cs))
        g = self._graphs.get(key)
        if g is None:
            g = self._capture(spec, noise, sizes, D)
            self._fill(g["host_in"], g["slots"], xs, ys, xqs, D)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                g["enqueue"]()                        # eager once: function attributes, allocator warm-up
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g["graph"] = None
            if self.use_graph:
                try:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        g["enqueue"]()
                    g["graph"] = graph
                except Exception as e:            # capture not possible here: stay concurrent, launch eagerly
                    import warnings
                    warnings.warn(f"CellBatch: CUDA graph capture failed ({type(e).__name__}: {e}); running the cells concurrently "
                                  "without a graph", RuntimeWarning)
                    self.use_graph = False
                    torch.cuda.synchronize(self.device)
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = g
        self._fill(g["host_in"], g["slots"], xs, ys, xqs, D)
        if g["graph"] is not None:
            g["graph"].replay()
            self.graph_replays += 1
        else:
            g["enqueue"]()
        torch.cuda.current_stream(self.device).synchronize()
        res = g["host_out"].numpy()
        means = res[:C * m].reshape(C, m).copy()
        vars_ = res[C * m:2 * C * m].reshape(C, m).copy()
        lmls = [float(v) for v in res[2 * C * m:2 * C * m + C]]
        info = g["host_info"].numpy()
        for ci in range(C):
            if int(info[ci]) != INT_MAX or not np.isfinite(lmls[ci]):
                # not positive definite at the first attempt (or NaN): the standard path owns the jitter retries / exceptions
                dev = self.device
                x, y, xq = (torch.tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev) for a in (xs[ci], ys[ci], xqs[ci]))
                st = E.fit(spec, x, y, noise, xq=xq)
                mean, var = E.predict(st, xq)
                means[ci], vars_[ci], lmls[ci] = mean.cpu().numpy(), var.cpu().numpy(), st.lml
        return means, vars_, lmls
