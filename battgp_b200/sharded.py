"""One exact GP sharded over the GPUs of a box (BASELINE.json configs[4]): block-row-cyclic Cholesky with
an NCCL exchange per panel.

Ownership: the N x N covariance is cut into NB-row stripes; stripe i lives on rank ``i % P`` as a ragged
``[nb_i, (i+1)*NB]`` matrix (only the columns up to the diagonal are ever stored: N = 200k is 20 GB per GPU on 8 GPUs).
Every rank builds its own stripes from the replicated X (32 N bytes) with the fused covariance kernel -- no traffic.

Per panel k (right-looking):
    owner(k)   : L_kk = chol(A_kk)                      bgp_potrf_block   (leaf kernels + DMMA GEMMs)
    all ranks  : broadcast L_kk and its 128-block inverses               NCCL broadcast  (NB^2 * 8 B)
    all ranks  : own rows of the panel  P_i <- A_ik L_kk^-T              bgp_trsm_rlt    (one call, rows packed)
    all ranks  : all-gather the packed panel                             NCCL all-gather ((N - k NB) * NB * 8 B)
    all ranks  : own stripes  A_ij -= P_i P_j^T  (k < j <= i)           bgp_gemm_nt     (tri mask on the diagonal block)
Solves: X <- W L^-T is right-looking over column blocks (owner solves its block with L_kk, broadcasts the M x NB result,
everyone updates its own column blocks); L^T alpha = z runs backwards with a reduce of the accumulated contributions.
LML pieces (sum log L_ii, z.z) and the predictive sums are all-reduced scalars / M-vectors.

The numerical work goes through an ``ops`` object: ``CudaOps`` (the engine -- the product) or, in the CPU tests only, an
injected checker that runs the same schedule over gloo.  Collectives are torch.distributed (NCCL over NVLink on GPUs).
"""
from __future__ import annotations

import json
import math
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import engine as E

INT_MAX = 2 ** 31 - 1


class CudaOps:
    """Block-level operations on one GPU through the C-ABI."""

    def __init__(self, device, spec: E.KernelSpec):
        self.device = torch.device(device)
        self.eng = E.get_engine(self.device)
        self.spec = spec

    def empty(self, rows, cols):
        return E.alloc_matrix(rows, cols, self.device)

    def zeros_vec(self, n):
        return torch.zeros(n, dtype=torch.float64, device=self.device)

    def scalars(self):
        return (torch.full((1,), INT_MAX, dtype=torch.int32, device=self.device),
                torch.zeros(1, dtype=torch.float64, device=self.device))

    def cov_block(self, x_rows, x_cols, out):
        self.eng.cov_build(self.spec, x_rows, x_cols, out=out)

    def cov_diag(self, x):
        return self.eng.cov_diag(self.spec, x)

    def potrf_block(self, A, info, logdet):
        return self.eng.potrf_block(A, info, logdet)

    def trsm_rlt(self, L, dinv, X):
        if self.eng.ozaki and X.shape[0] >= 1024 and L.shape[0] >= 1024:
            # scratch for the int8 path of the big TRSM updates: digit planes of X[:, :n/2] and L21
            k = L.shape[0] // 2
            self.eng.ensure_workspace_bytes(int(self.eng.L.bgp_oz_slice_bytes(X.shape[0], k)) +
                                            int(self.eng.L.bgp_oz_slice_bytes(k, k)) + 4096)
        self.eng.trsm_rlt(L, dinv, X)

    def gemm_nt(self, A, B, C, alpha, beta, tri=False, roff=0, coff=0):
        self.eng.gemm_nt(A, B, C, alpha=alpha, beta=beta, tri=tri, roff=roff, coff=coff)

    @property
    def has_oz(self):
        return self.eng.ozaki

    def oz_slice(self, P, buf):
        return self.eng.oz_slice(P, buf)

    def oz_slice_gather(self, P, rows, blkmap, blkrows, buf):
        return self.eng.oz_slice_gather(P, rows, blkmap, blkrows, buf)

    def oz_gemm(self, buf, rows, arow0, brow0, C, K, alpha, tri, roff, coff):
        self.eng.oz_gemm(buf, rows, arow0, buf, rows, brow0, C, K, alpha=alpha, tri=tri, roff=roff, coff=coff)

    def trsv(self, L, dinv, b, trans):
        self.eng.trsv(L, dinv, b, trans)

    def gemv_t(self, A, v, y, alpha):
        self.eng.gemv_t(A, v, y, alpha)

    def rowsumsq(self, V, out, accumulate):
        self.eng.rowsumsq(V, out, accumulate)


class ShardedGP:
    def __init__(self, spec: E.KernelSpec, x: torch.Tensor, y: torch.Tensor, noise: float, *, nb: int = 1024,
                 group=None, ops=None):
        if nb % 128:
            raise ValueError("nb must be a multiple of 128")
        self.spec, self.x, self.y, self.noise, self.NB = spec, x.contiguous(), y.contiguous(), float(noise), nb
        self.group = group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ops = ops if ops is not None else CudaOps(x.device, spec)
        self.N = x.shape[0]
        self.nblk = (self.N + nb - 1) // nb
        self.owned = [i for i in range(self.nblk) if i % self.P == self.rank]
        self.rows: Dict[int, torch.Tensor] = {}
        self.dinv: Dict[int, torch.Tensor] = {}
        self.alpha: Optional[torch.Tensor] = None
        self.bytes_received = 0
        self.profile = False            # True: synchronise after every phase and accumulate seconds in self.phase_s
        self.phase_s: Dict[str, float] = {}

    def _tick(self, name, t0):
        if not self.profile:
            return t0
        if self.x.is_cuda:
            torch.cuda.synchronize(self.x.device)
        t1 = time.perf_counter()
        self.phase_s[name] = self.phase_s.get(name, 0.0) + (t1 - t0)
        return t1

    # ---- geometry
    def b0(self, i):
        return i * self.NB

    def e(self, i):
        return min(self.N, (i + 1) * self.NB)

    def nbi(self, i):
        return self.e(i) - self.b0(i)

    def _global_rank(self, r):
        return dist.get_global_rank(self.group, r) if (self.group is not None and dist.is_initialized()) else r

    def _bcast(self, t, owner):
        if self.P > 1:
            dist.broadcast(t, src=self._global_rank(owner), group=self.group)
            if owner != self.rank:
                self.bytes_received += t.numel() * 8

    def _allreduce(self, t, op=None):
        if self.P > 1:
            dist.all_reduce(t, op=op if op is not None else dist.ReduceOp.SUM, group=self.group)

    # ---- fit
    def build(self):
        for i in self.owned:
            b0, e = self.b0(i), self.e(i)
            blk = self.ops.empty(e - b0, e)
            self.ops.cov_block(self.x[b0:e], self.x[:e], blk)
            blk[:, b0:e].diagonal().add_(self.noise)
            self.rows[i] = blk

    def factor(self):
        ops, P, NB = self.ops, self.P, self.NB
        info, logdet = ops.scalars()
        t0 = time.perf_counter()
        for k in range(self.nblk):
            owner = k % P
            b0k, ek, nbk = self.b0(k), self.e(k), self.nbi(k)
            Lkk = ops.empty(nbk, nbk)
            ndinv = ((nbk + 127) // 128) * 128 * 128
            if self.rank == owner:
                Akk = self.rows[k][:, b0k:ek]
                dkk = ops.potrf_block(Akk, info, logdet)
                self.dinv[k] = dkk
                Lkk.copy_(Akk)
            else:
                dkk = ops.zeros_vec(ndinv)
            if k == self.nblk - 1:
                break
            t0 = self._tick("diag_potrf", t0)
            self._bcast(Lkk, owner)
            self._bcast(dkk, owner)
            t0 = self._tick("bcast", t0)
            mine = [i for i in self.owned if i > k]
            nrows = sum(self.nbi(i) for i in mine)
            # pack this rank's rows of the panel, solve them in one call, scatter back (they are part of L)
            nbelow = self.nblk - 1 - k
            cnt_max = (nbelow + P - 1) // P
            send = ops.empty(cnt_max * NB, nbk)
            if nrows:
                o = 0
                for i in mine:
                    send[o:o + self.nbi(i)].copy_(self.rows[i][:, b0k:ek])
                    o += NB
                if self.nbi(mine[-1]) < NB:            # ragged last stripe: keep the padding finite
                    send[o - NB + self.nbi(mine[-1]):o].zero_()
                ops.trsm_rlt(Lkk, dkk, send[:(len(mine) - 1) * NB + self.nbi(mine[-1])])
                o = 0
                for i in mine:
                    self.rows[i][:, b0k:ek].copy_(send[o:o + self.nbi(i)])
                    o += NB
            t0 = self._tick("panel_trsm", t0)
            # exchange: every rank ends up with the whole panel.  The all-gather result is rank-major; the int8 path
            # slices it straight into stripe order (block map), the DMMA path needs a reordered fp64 copy.
            use_oz = getattr(ops, "has_oz", False) and nbk % 64 == 0 and len(mine) > 0
            panel = None
            panel_rows = nbelow * NB
            if P > 1:
                assert send.is_contiguous()
                recv = torch.empty((P * cnt_max * NB, nbk), dtype=send.dtype, device=send.device)
                dist.all_gather_into_tensor(recv, send, group=self.group)
                self.bytes_received += (P - 1) * cnt_max * NB * nbk * 8
                first = {r: next(j for j in range(k + 1, k + 1 + P) if j % P == r) for r in range(P)}
                idx = [(j % P) * cnt_max + (j - first[j % P]) // P for j in range(k + 1, self.nblk)]
                t0 = self._tick("allgather", t0)
                if not mine:
                    pass                                     # this rank has no stripe below the panel: nothing to update
                elif use_oz and hasattr(ops, "oz_slice_gather"):
                    blkmap = torch.tensor(idx, dtype=torch.int32, device=send.device)
                    self._ozbuf = ops.oz_slice_gather(recv, panel_rows, blkmap, NB, getattr(self, "_ozbuf", None))
                else:
                    sel = torch.tensor(idx, dtype=torch.long, device=send.device)
                    panel = recv.view(P * cnt_max, NB, nbk).index_select(0, sel).view(-1, nbk)
                    if use_oz:
                        self._ozbuf = ops.oz_slice(panel, getattr(self, "_ozbuf", None))
                del recv
            else:
                panel = send
                if use_oz:
                    self._ozbuf = ops.oz_slice(panel, getattr(self, "_ozbuf", None))
            t0 = self._tick("reorder+slice", t0)
            for t, i in enumerate(mine):
                b0i, ei = self.b0(i), self.e(i)
                C = self.rows[i][:, ek:ei]
                if use_oz:
                    ops.oz_gemm(self._ozbuf, panel_rows if P > 1 else panel.shape[0], (i - k - 1) * NB, 0, C, nbk, -1.0, True, b0i, ek)
                else:
                    A = panel[(i - k - 1) * NB:(i - k - 1) * NB + self.nbi(i)]
                    B = panel[:ei - ek]
                    ops.gemm_nt(A, B, C, -1.0, 1.0, tri=True, roff=b0i, coff=ek)
            t0 = self._tick("trailing_update", t0)
        self._allreduce(logdet)
        self._allreduce(info, dist.ReduceOp.MIN if dist.is_initialized() else None)
        self.logdet = float(logdet.item())
        inf = int(info.item())
        self.info = 0 if inf == INT_MAX else inf
        return self.info

    def solve_rlt(self, Wb: Dict[int, torch.Tensor], m: int):
        """In place X <- W L^-T for W given as owned column blocks {i: [m, nb_i]}."""
        ops, P = self.ops, self.P
        for k in range(self.nblk):
            owner = k % P
            b0k, ek, nbk = self.b0(k), self.e(k), self.nbi(k)
            Xk = ops.empty(m, nbk)
            if self.rank == owner:
                ops.trsm_rlt(self.rows[k][:, b0k:ek], self.dinv[k], Wb[k])
                Xk.copy_(Wb[k])
            if k == self.nblk - 1:
                break
            self._bcast(Xk, owner)          # nb_k is even for every block that has stripes below it -> contiguous
            for i in self.owned:
                if i > k:
                    ops.gemm_nt(Xk, self.rows[i][:, b0k:ek], Wb[i], -1.0, 1.0)
        return Wb

    def solve_alpha(self):
        """z = L^-1 y (forward, as a 1-row solve_rlt), alpha = L^-T z (backward with reduced contributions)."""
        ops, P = self.ops, self.P
        zb = {}
        for i in self.owned:
            w = ops.empty(1, self.nbi(i))
            w[0].copy_(self.y[self.b0(i):self.e(i)])
            zb[i] = w
        self.solve_rlt(zb, 1)
        zz = ops.zeros_vec(1)
        for i in self.owned:
            ops.rowsumsq(zb[i], zz, True)
        self._allreduce(zz)
        self.zz = float(zz.item())
        acc = ops.zeros_vec(self.N)
        alpha = ops.zeros_vec(self.N)
        for k in reversed(range(self.nblk)):
            owner = k % P
            b0k, ek = self.b0(k), self.e(k)
            tot = acc[b0k:ek].clone()
            if P > 1:
                dist.reduce(tot, dst=self._global_rank(owner), op=dist.ReduceOp.SUM, group=self.group)
            if self.rank == owner:
                rhs = zb[k][0].clone() - tot
                ops.trsv(self.rows[k][:, b0k:ek], self.dinv[k], rhs, True)
                alpha[b0k:ek].copy_(rhs)
                if b0k > 0:
                    ops.gemv_t(self.rows[k][:, :b0k], rhs, acc[:b0k], 1.0)
        self._allreduce(alpha)
        self.alpha = alpha
        n = self.N
        self.lml = -0.5 * self.zz - 0.5 * self.logdet - 0.5 * n * math.log(2.0 * math.pi)
        return alpha

    def fit(self):
        t0 = time.perf_counter()
        self.build()
        t0 = self._tick("build", t0)
        info = self.factor()
        if info != 0:
            raise E.NotPSDError(f"sharded Cholesky failed at pivot {info}")
        t0 = time.perf_counter()
        self.solve_alpha()
        self._tick("solve_alpha", t0)
        return self

    def residual(self) -> float:
        """Matrix-free check at any size: || K alpha - y || / || y || with the rows of K REBUILT from X (full width)."""
        ops = self.ops
        rr = ops.zeros_vec(1)
        a_row = ops.empty(1, self.N)
        a_row[0].copy_(self.alpha)
        for i in self.owned:
            b0, e = self.b0(i), self.e(i)
            krow = ops.empty(e - b0, self.N)
            ops.cov_block(self.x[b0:e], self.x, krow)
            r = ops.empty(e - b0, 1)
            r[:, 0].copy_(self.noise * self.alpha[b0:e] - self.y[b0:e])
            ops.gemm_nt(krow, a_row, r, 1.0, 1.0)
            rt = ops.empty(1, e - b0)
            rt[0].copy_(r[:, 0])
            ops.rowsumsq(rt, rr, True)
            del krow
        self._allreduce(rr)
        return math.sqrt(float(rr.item())) / float(torch.linalg.vector_norm(self.y).item())

    # ---- predict
    def predict(self, xq: torch.Tensor, clamp: bool = True):
        ops = self.ops
        m = xq.shape[0]
        xq = xq.contiguous()
        Wb = {}
        mean = ops.empty(m, 1)
        mean.zero_()
        for i in self.owned:
            b0, e = self.b0(i), self.e(i)
            w = ops.empty(m, e - b0)
            ops.cov_block(xq, self.x[b0:e], w)
            Wb[i] = w
            a = ops.empty(1, e - b0)
            a[0].copy_(self.alpha[b0:e])
            ops.gemm_nt(w, a, mean, 1.0, 1.0)
        mean = mean[:, 0].contiguous()
        self._allreduce(mean)
        self.solve_rlt(Wb, m)
        ss = ops.zeros_vec(m)
        for i in self.owned:
            ops.rowsumsq(Wb[i], ss, True)
        self._allreduce(ss)
        var = ops.cov_diag(xq) - ss
        if clamp:
            var = var.clamp_min(E.MIN_VARIANCE_F64)
        return mean, var


# ------------------------------------------------------------------------------------------------------- bench
def bench(args, rank: int, world: int, dev: torch.device):
    """bench.py --workload sharded: BASELINE configs[4] (one GP over all ranks)."""
    from .synth import query_grid, synth_field_data
    import bench as B                                    # the repo-root bench.py (flops formula, clock sampler)

    n = args.n
    x_np, y_np = synth_field_data(n, seed=0)
    xq_np = query_grid(x_np, B.M_QUERY)
    x, y, xq = (torch.tensor(a, device=dev) for a in (x_np, y_np, xq_np))
    spec = E.battgp_spec()
    eng = E.get_engine(dev)

    def step(profile=False):
        gp = ShardedGP(spec, x, y, B.NOISE, nb=args.nb)
        gp.profile = profile
        gp.fit()
        t0 = time.perf_counter()
        mean, var = gp.predict(xq)
        gp._tick("predict", t0)
        return gp, mean, var

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        gp, mean, var = step()
        del gp
        torch.cuda.empty_cache()
    sampler = B.ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    recv = 0
    for it in range(args.steps):
        gp, mean, var = step()
        recv = gp.bytes_received
        lml = gp.lml
        if it < args.steps - 1:
            del gp
    e1.record()
    sync()
    resid = gp.residual() if args.verify else None
    del gp
    phases = None
    if args.phases:
        torch.cuda.empty_cache()
        gpp, _, _ = step(profile=True)
        phases = {k: round(v, 4) for k, v in gpp.phase_s.items()}
        del gpp
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        sec = ms / 1e3 / args.steps
        flops = B.algorithmic_flops(n)
        line = {"metric": "exact_gp_fit_predict_gflops", "value": flops / sec * 1e-9, "unit": "GF/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": B.workload_name(args, world), "n": n, "m_query": B.M_QUERY, "kernel": "wiener+rbf_ard",
                           "nb": args.nb, "fit_predict_seconds": sec, "lml": lml, "mean0": float(mean[0]), "var0": float(var[0]),
                           "nccl_bytes_received_per_rank_per_step": recv,
                           "residual_Kalpha_minus_y_over_y": resid, "phase_seconds_rank0_synchronised": phases,
                           "l2_policy": "inputs_exceed_l2 (per-rank stripes rebuilt every step)"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": flops / sec * 1e-9, "unit": "GF/s", "note": "X,y replicated in HBM; host e2e measured on the per_gpu workload",
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 2 * B.M_QUERY * 8},
                "roofline": B.step_roofline(n, sec, world, eng.ozaki)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
