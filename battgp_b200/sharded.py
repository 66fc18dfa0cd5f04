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

    def oz_gemm(self, buf, rows, arow0, brow0, C, K, alpha, tri, roff, coff, tpc=0):
        self.eng.oz_gemm(buf, rows, arow0, buf, rows, brow0, C, K, alpha=alpha, tri=tri, roff=roff, coff=coff, tpc=tpc)

    def oz_gemm_ab(self, bufA, rowsA, bufB, rowsB, brow0, C, K, alpha):
        """C += alpha A B^T with A and B from two different slice buffers (A from row 0, full rectangle)."""
        self.eng.oz_gemm(bufA, rowsA, 0, bufB, rowsB, brow0, C, K, alpha=alpha, tri=False, roff=0, coff=0, tpc=0)

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
        # extra rows riding through the factorisation (fit(xq=...)): E = [K(xq, X); y^T], (m + 1) x N, owned by ONE rank; they
        # take part in every panel solve and trailing update and come out as [V = K_*N L^-T; z^T = (L^-1 y)^T] -- the
        # sharded form of bgp_potrf_aug: no separate 98-step solve chain for the predictive variance and for z
        self.E: Optional[torch.Tensor] = None
        self.e_owner = 0
        self.xq_aug: Optional[torch.Tensor] = None
        self.bytes_received = 0
        import os as _os
        # one panel of look-ahead on a side stream (GPU only).  OFF by default: measured on 8 x B200 at N = 200 000 it is
        # slower than the serial schedule (4.61 s against 4.48 s; 4.98 / 5.22 s with 16 / 64 tiles per CTA,
        # profiles/probe_r02_sharded_8gpu_overlap.jsonl) -- the updates have to leave the persistent form so that the side
        # stream finds SMs, which costs more than the 0.35 s of exchange and diagonal-block latency it hides.
        self.lookahead = _os.environ.get("BATTGP_SHARDED_LOOKAHEAD", "0") != "0"
        self.tpc_long = int(_os.environ.get("BATTGP_SHARDED_TPC", "4"))     # tiles per CTA of the long trailing updates while the side stream needs SMs
        self.tpc_short = int(_os.environ.get("BATTGP_SHARDED_TPC_SHORT", "1"))
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
        if self.xq_aug is not None:
            self.e_owner = self.nblk % self.P                      # the rank that owns the fewest stripes
            if self.rank == self.e_owner:
                m = self.xq_aug.shape[0]
                self.E = self.ops.empty(m + 1, self.N)
                self.ops.cov_block(self.xq_aug, self.x, self.E[:m])
                self.E[m].copy_(self.y)

    # ---- streams (no-ops for the CPU checker of tests/test_sharded_cpu.py)
    def _cuda(self):
        return self.x.is_cuda

    def _mk_streams(self):
        if not self._cuda():
            return None, None
        main = torch.cuda.current_stream(self.x.device)
        lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
        return main, torch.cuda.Stream(self.x.device, priority=hi)

    class _On:
        """``with _On(stream):`` -- torch.cuda.stream(stream) on GPUs, nothing on CPU."""

        def __init__(self, stream):
            self.cm = torch.cuda.stream(stream) if stream is not None else None

        def __enter__(self):
            if self.cm is not None:
                self.cm.__enter__()

        def __exit__(self, *a):
            if self.cm is not None:
                self.cm.__exit__(*a)

    def _panel_geometry(self, k):
        """Who holds which rows of panel k in the packed send / rank-major gather buffers."""
        P, NB = self.P, self.NB
        nbelow = self.nblk - 1 - k
        cnt_max = (nbelow + P - 1) // P
        first = {r: next(j for j in range(k + 1, k + 1 + P) if j % P == r) for r in range(P)}
        idx = [(j % P) * cnt_max + (j - first[j % P]) // P for j in range(k + 1, self.nblk)]
        return nbelow, cnt_max, idx

    def _update(self, k, i, c0, c1, ozbuf, panel, panel_rows, tpc=0):
        """rows[i][:, c0:c1] -= P_i P_j^T for the columns [c0, c1) (global indices, c0 >= e_k) with panel k."""
        if c1 <= c0:
            return
        ops, NB = self.ops, self.NB
        ek, nbk = self.e(k), self.nbi(k)
        b0i = self.b0(i)
        C = self.rows[i][:, c0:c1]
        if ozbuf is not None:
            ops.oz_gemm(ozbuf, panel_rows, (i - k - 1) * NB, c0 - ek, C, nbk, -1.0, True, b0i, c0, tpc)
        else:
            A = panel[(i - k - 1) * NB:(i - k - 1) * NB + self.nbi(i)]
            B = panel[c0 - ek:c1 - ek]
            ops.gemm_nt(A, B, C, -1.0, 1.0, tri=True, roff=b0i, coff=c0)

    def factor(self):
        """Right-looking over the NB-wide panels with ONE PANEL OF LOOK-AHEAD (the structure csrc/api.cu potrf_driver uses
        on one GPU): the update of column block k+1 with panel k, the factorisation of A_k+1,k+1, its broadcast, the panel
        solve, the all-gather and the int8 slicing of panel k+1 all run on a high-priority side stream (NCCL collectives are
        enqueued on the stream that is current) while the main stream still applies panel k to the column blocks >= k+2.
        ``self.profile`` (per-phase seconds) synchronises after every phase and therefore runs the same schedule serially."""
        ops, P, NB = self.ops, self.P, self.NB
        info, logdet = ops.scalars()
        nblk = self.nblk
        main, side = self._mk_streams()
        overlap = side is not None and not self.profile and getattr(self, "lookahead", True) and self.xq_aug is None
        if not overlap:
            side = None
        use_oz_all = getattr(ops, "has_oz", False)
        # persistent exchange buffers, double-buffered by panel parity (no allocator traffic across streams)
        _, cnt0, _ = self._panel_geometry(0) if nblk > 1 else (0, 0, [])
        nbk0 = self.nbi(0)
        send_buf = [ops.empty(max(cnt0, 1) * NB, nbk0) for _ in range(2)]
        recv_buf = [torch.empty((P * max(cnt0, 1) * NB, nbk0), dtype=send_buf[0].dtype, device=send_buf[0].device) for _ in range(2)] if P > 1 else None
        ozbufs = [None, None]
        ev_panel = [None, None]          # panel k gathered + sliced (side)
        ev_u2a = [None, None]            # column block k+2 brought up to date with panel k (main)
        ev_u2 = [None, None]             # every update with panel k issued (main): its buffers may be overwritten
        state = {}
        state_E = {}
        aug = self.xq_aug is not None    # extra rows ride along (all ranks must then take part in the LAST panel's broadcast too)
        ozE = None

        def record(stream):
            if stream is None:
                return None
            e = torch.cuda.Event()
            e.record(stream)
            return e

        def wait(stream, e):
            if stream is not None and e is not None:
                stream.wait_event(e)

        def panel_chain(k):
            """Column block k is up to date: factor A_kk, broadcast, solve the rows below, exchange, slice.  Runs on the
            stream that is current."""
            t0 = time.perf_counter()
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
            if k == nblk - 1 and not aug:
                return
            t0 = self._tick("diag_potrf", t0)
            self._bcast(Lkk, owner)
            self._bcast(dkk, owner)
            t0 = self._tick("bcast", t0)
            if self.E is not None:                                 # this rank owns the extra rows: solve their block of panel k
                Eb = ops.empty(self.E.shape[0], nbk)
                Eb.copy_(self.E[:, b0k:ek])
                ops.trsm_rlt(Lkk, dkk, Eb)
                self.E[:, b0k:ek].copy_(Eb)
                state_E[k] = Eb
            if k == nblk - 1:
                return
            mine = [i for i in self.owned if i > k]
            nrows = sum(self.nbi(i) for i in mine)
            nbelow, cnt_max, idx = self._panel_geometry(k)
            b = k & 1
            # pack this rank's rows of the panel, solve them in one call, scatter back (they are part of L)
            send = send_buf[b][:cnt_max * NB, :nbk]
            if send.stride(0) != nbk:                       # narrower last panels: keep the exchange buffer contiguous
                send = send_buf[b].view(-1)[:cnt_max * NB * nbk].view(cnt_max * NB, nbk)
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
            need_panel = len(mine) > 0 or self.E is not None      # this rank updates something with panel k
            use_oz = use_oz_all and nbk % 64 == 0 and need_panel
            panel = None
            panel_rows = nbelow * NB
            if P > 1:
                recv = recv_buf[b].view(-1)[:P * cnt_max * NB * nbk].view(P * cnt_max * NB, nbk)
                dist.all_gather_into_tensor(recv, send, group=self.group)
                self.bytes_received += (P - 1) * cnt_max * NB * nbk * 8
                t0 = self._tick("allgather", t0)
                if not need_panel:
                    pass                                     # this rank has no stripe below the panel: nothing to update
                elif use_oz and hasattr(ops, "oz_slice_gather"):
                    ozbufs[b] = ops.oz_slice_gather(recv, panel_rows, self._blkmap(k, idx, recv.device), NB, ozbufs[b])
                else:
                    sel = torch.tensor(idx, dtype=torch.long, device=send.device)
                    panel = recv.view(P * cnt_max, NB, nbk).index_select(0, sel).view(-1, nbk)
                    if main is not None and side is not None:
                        panel.record_stream(main)          # allocated on the side stream, read by the main stream's updates
                    if use_oz:
                        ozbufs[b] = ops.oz_slice(panel, ozbufs[b])
            else:
                panel = send
                panel_rows = panel.shape[0]
                if use_oz:
                    ozbufs[b] = ops.oz_slice(panel, ozbufs[b])
            self._tick("reorder+slice", t0)
            state[k] = (mine, ozbufs[b] if use_oz else None, panel, panel_rows)

        with self._On(side):
            wait(side, record(main))                 # the build ran on the main stream
            panel_chain(0)
            ev_panel[0] = record(side)
        for k in range(nblk - 1):
            mine, ozb, panel, panel_rows = state.pop(k)
            ek, ek1 = self.e(k), self.e(k + 1)
            ek2 = self.e(k + 2) if k + 2 < nblk else ek1
            if self.E is not None:
                # extra rows first (the chain of panel k+1 below solves their next column block): E[:, e_k:] -= P_E P_j^T for
                # every row j below the panel.  Never concurrent with the look-ahead (it is off when rows ride along).
                t0 = time.perf_counter()
                Eb = state_E.pop(k)
                if ozb is not None:
                    ozE = ops.oz_slice(Eb, ozE)
                    ops.oz_gemm_ab(ozE, Eb.shape[0], ozb, panel_rows, 0, self.E[:, ek:], self.nbi(k), -1.0)
                else:
                    ops.gemm_nt(Eb, panel[:self.N - ek], self.E[:, ek:], -1.0, 1.0)
                self._tick("trailing_update", t0)
            # ---- side stream: column block k+1 <- panel k, then the whole chain of panel k+1
            with self._On(side):
                t0 = time.perf_counter()
                wait(side, ev_u2a[(k - 1) & 1] if k >= 1 else None)      # block k+1 has received panel k-1 (main)
                wait(side, ev_u2[(k - 1) & 1] if k >= 1 else None)       # buffers of parity (k+1)&1 are free again
                for i in mine:
                    self._update(k, i, ek, min(ek1, self.e(i)), ozb, panel, panel_rows, self.tpc_short if overlap else 0)
                self._tick("trailing_update", t0)
                panel_chain(k + 1)
                ev_panel[(k + 1) & 1] = record(side)
            # ---- main stream: the rest of the trailing update with panel k
            t0 = time.perf_counter()
            wait(main, ev_panel[k & 1])
            for i in mine:
                self._update(k, i, ek1, min(ek2, self.e(i)), ozb, panel, panel_rows, self.tpc_short if overlap else 0)
            ev_u2a[k & 1] = record(main)
            for i in mine:
                self._update(k, i, ek2, self.e(i), ozb, panel, panel_rows, self.tpc_long if overlap else 0)
            ev_u2[k & 1] = record(main)
            self._tick("trailing_update", t0)
            del panel
        if side is not None:
            main.wait_stream(side)
        self._allreduce(logdet)
        self._allreduce(info, dist.ReduceOp.MIN if dist.is_initialized() else None)
        self.logdet = float(logdet.item())
        inf = int(info.item())
        self.info = 0 if inf == INT_MAX else inf
        return self.info

    def _blkmap(self, k, idx, device):
        """int32 device map (logical stripe of panel k -> block of the rank-major gather buffer); built once per panel on
        the host and uploaded from pinned memory without a synchronising copy."""
        t = torch.tensor(idx, dtype=torch.int32)
        if device.type == "cuda":
            t = t.pin_memory().to(device, non_blocking=True)
        return t

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
        if self.xq_aug is not None:
            # z^T left the factorisation as the last of the extra rows: one broadcast instead of the forward chain
            z = ops.zeros_vec(self.N)
            if self.E is not None:
                z.copy_(self.E[self.E.shape[0] - 1])
            self._bcast(z, self.e_owner)
            for i in self.owned:
                w = ops.empty(1, self.nbi(i))
                w[0].copy_(z[self.b0(i):self.e(i)])
                zb[i] = w
            zrow = ops.empty(1, self.N)
            zrow[0].copy_(z)
            zz = ops.zeros_vec(1)
            ops.rowsumsq(zrow, zz, False)
        else:
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

    def fit(self, xq: Optional[torch.Tensor] = None):
        """``xq`` (M query points): K(xq, X) and y ride through the factorisation as extra rows (the sharded bgp_potrf_aug), so
        that predict(xq) needs no solve chain of its own -- BattGP fits a model and predicts once on one grid
        (battgp_full.py:98-120)."""
        t0 = time.perf_counter()
        self.xq_aug = None if xq is None else xq.contiguous()
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
        ss = ops.zeros_vec(m)
        if self.xq_aug is not None and self.xq_aug.shape == xq.shape and torch.equal(self.xq_aug, xq):
            if self.E is not None:                      # V = K_*N L^-T left the factorisation with the extra rows
                ops.rowsumsq(self.E[:m], ss, False)
            self._bcast(ss, self.e_owner)
        else:
            self.solve_rlt(Wb, m)
            for i in self.owned:
                ops.rowsumsq(Wb[i], ss, True)
            self._allreduce(ss)
        var = ops.cov_diag(xq) - ss
        if clamp:
            var = var.clamp_min(E.MIN_VARIANCE_F64)
        return mean, var


# ------------------------------------------------------------------------------------------------------- bench
def _sync(dev, world):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)


def _timed_steps(n, nb, rank, world, dev, steps, warmup, verify, phases, e2e=False):
    """`steps` timed fit+predict passes of ONE GP of size n over all ranks (max over ranks, CUDA events)."""
    from .synth import query_grid, synth_field_data
    import bench as B                                    # the repo-root bench.py (flops formula, clock sampler)
    x_np, y_np = synth_field_data(n, seed=0)
    xq_np = query_grid(x_np, B.M_QUERY)
    x, y, xq = (torch.tensor(a, device=dev) for a in (x_np, y_np, xq_np))
    xh, yh, xqh = (torch.tensor(a).pin_memory() for a in (x_np, y_np, xq_np))
    spec = E.battgp_spec()
    eng = E.get_engine(dev)

    def step(profile=False, host=False):
        if host:                                        # e2e: every rank stages the replicated inputs from pinned host memory
            xs, ys, xqs = xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True), xqh.to(dev, non_blocking=True)
        else:
            xs, ys, xqs = x, y, xq
        gp = ShardedGP(spec, xs, ys, B.NOISE, nb=nb)
        gp.profile = profile
        gp.fit(xqs)
        t0 = time.perf_counter()
        mean, var = gp.predict(xqs)
        gp._tick("predict", t0)
        if host:
            mean, var = mean.cpu(), var.cpu()
        return gp, mean, var

    def timed(nsteps, host=False):
        _sync(dev, world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for it in range(nsteps):
            out = None
            out = step(host=host)
            if it < nsteps - 1:
                out = None
        e1.record()
        _sync(dev, world)
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(warmup):
        step()
        torch.cuda.empty_cache()
    sampler = B.ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ms, (gp, mean, var) = timed(steps)
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    res = {"n": n, "nb": nb, "steps": steps, "warmup": warmup, "seconds_per_step": ms / 1e3 / steps, "lml": gp.lml,
           "mean0": float(mean[0]), "var0": float(var[0]), "nccl_bytes_received_per_rank_per_step": gp.bytes_received,
           "gpu_launches": int(launches), "clocks": clocks,
           "predictions_finite_and_positive": bool(torch.isfinite(mean).all() and torch.isfinite(var).all() and (var > 0).all())}
    res["residual_Kalpha_minus_y_over_y"] = gp.residual() if verify else None
    del gp
    torch.cuda.empty_cache()
    if e2e:
        ms_h, _ = timed(max(1, min(steps, 2)), host=True)
        res["e2e_seconds_per_step"] = ms_h / 1e3 / max(1, min(steps, 2))
        res["h2d_bytes_per_step"] = int(xh.numel() + yh.numel() + xqh.numel()) * 8
    if phases:
        torch.cuda.empty_cache()
        gpp, _, _ = step(profile=True)
        res["phase_seconds_rank0_serialised"] = {k: round(v, 4) for k, v in gpp.phase_s.items()}
        del gpp
        torch.cuda.empty_cache()
    return res


def parity(rank, world, dev, checker=None):
    """Multi-rank parity, run inside the bench so that the driver's N>1 runs prove it (SURVEY 8c iii):
    (1) n=5000 against ``checker(x, y, xq) -> (mean, var, lml)`` -- bench.py injects the CPU oracle here (this package never
        imports it): mean rtol 1e-7, variance 1e-6, LML 1e-9;
    (2) N=40 000: the sharded result over `world` ranks against the single-GPU engine (bgp_potrf_aug) on the same box."""
    import numpy as np
    from .synth import query_grid, synth_field_data
    import bench as B
    out = {}
    spec = E.battgp_spec()
    # (1) oracle
    n1 = 5000
    x_np, y_np = synth_field_data(n1, seed=7)
    xq_np = query_grid(x_np, 64)
    gp = ShardedGP(spec, torch.tensor(x_np, device=dev), torch.tensor(y_np, device=dev), B.NOISE, nb=256)
    xq1 = torch.tensor(xq_np, device=dev)
    gp.fit(xq1)
    mean, var = gp.predict(xq1)
    if rank == 0 and checker is not None:
        mr, vr, lml_ref = checker(x_np, y_np, xq_np)
        o = {"n": n1, "nb": 256, "mean_max_rel": float(np.max(np.abs(mean.cpu().numpy() - mr) / np.abs(mr))),
             "var_max_rel": float(np.max(np.abs(var.cpu().numpy() - vr) / np.abs(vr))),
             "lml_rel": abs(gp.lml - lml_ref) / abs(lml_ref), "tolerance": {"mean": 1e-7, "var": 1e-6, "lml": 1e-9}}
        o["ok"] = bool(o["mean_max_rel"] < 1e-7 and o["var_max_rel"] < 1e-6 and o["lml_rel"] < 1e-9)
        out["vs_cpu_oracle"] = o
    del gp
    # (2) rank-count invariance at N = 40 000
    n2 = 40000
    x_np, y_np = synth_field_data(n2, seed=0)
    xq_np = query_grid(x_np, B.M_QUERY)
    xd, yd, xqd = (torch.tensor(a, device=dev) for a in (x_np, y_np, xq_np))
    gp = ShardedGP(spec, xd, yd, B.NOISE, nb=1024)
    gp.fit(xqd)
    mean, var = gp.predict(xqd)
    lml = gp.lml
    del gp
    torch.cuda.empty_cache()
    if rank == 0:
        st = E.fit(spec, xd, yd, B.NOISE, xq=xqd)
        m1, v1 = E.predict(st, xqd)
        o = {"n": n2, "ranks": world, "mean_max_rel": float(((mean - m1).abs() / m1.abs()).max()),
             "var_max_rel": float(((var - v1).abs() / v1.abs()).max()), "lml_rel": abs(lml - st.lml) / abs(st.lml),
             "alpha_note": "compared through mean/variance/LML", "tolerance": {"mean": 1e-6, "var": 1e-4, "lml": 1e-9}}
        o["ok"] = bool(o["mean_max_rel"] < 1e-6 and o["var_max_rel"] < 1e-4 and o["lml_rel"] < 1e-9)
        out["vs_single_gpu_engine"] = o
        del st
        E.get_engine(dev).release_workspace()
    torch.cuda.empty_cache()
    _sync(dev, world)
    return out


def bench_object(args, rank: int, world: int, dev: torch.device, checker=None):
    """The "sharded" object of bench.py's N>1 line: BASELINE configs[4] timed in the same run + its parity numbers."""
    import bench as B
    n = args.sharded_n
    par = parity(rank, world, dev, checker)
    r = _timed_steps(n, 2048 if n >= 100000 else 1024, rank, world, dev, args.sharded_steps, 1, True, True)
    if rank != 0:
        return None
    sec = r["seconds_per_step"]
    flops = B.algorithmic_flops(n)
    eng = E.get_engine(dev)
    return {"config": {"workload": f"full_gp Wiener+RBF-ARD N={n} block-row-sharded Cholesky over {world} GPU(s), NCCL panel broadcast + "
                                   f"all-gather, query rows and y riding through the factorisation (BASELINE configs[4])",
                       "n": n, "m_query": B.M_QUERY, "kernel": "wiener+rbf_ard", "nb": r["nb"]},
            "metric": B.METRIC, "value": flops / sec * 1e-9, "unit": "GF/s", "scaling": "strong", "steps": r["steps"], "warmup": r["warmup"],
            "ms_per_step": sec * 1e3, "fit_predict_seconds": sec, "timing": "CUDA events, max over ranks, barrier on both sides",
            "nccl_bytes_received_per_rank_per_step": r["nccl_bytes_received_per_rank_per_step"],
            "phase_seconds_rank0_serialised": r.get("phase_seconds_rank0_serialised"),
            "residual_Kalpha_minus_y_over_y": r["residual_Kalpha_minus_y_over_y"], "lml": r["lml"],
            "predictions_finite_and_positive": r["predictions_finite_and_positive"],
            "gpu_launches_per_rank": r["gpu_launches"], "parity": par, "roofline": B.step_roofline(n, sec, world, eng.ozaki)}


def bench(args, rank: int, world: int, dev: torch.device, checker=None):
    """bench.py --workload sharded: BASELINE configs[4] (one GP over all ranks) as the primary line."""
    import bench as B
    n = args.n
    r = _timed_steps(n, args.nb, rank, world, dev, args.steps, args.warmup, args.verify, args.phases, e2e=True)
    par = parity(rank, world, dev, checker) if (args.verify and world > 1) else None
    if rank != 0:
        return None
    sec = r["seconds_per_step"]
    flops = B.algorithmic_flops(n)
    eng = E.get_engine(dev)
    return {"metric": B.METRIC, "value": flops / sec * 1e-9, "unit": "GF/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": B.workload_config(args, world),
            "detail": {"nb": args.nb, "fit_predict_seconds": sec, "lml": r["lml"], "mean0": r["mean0"], "var0": r["var0"],
                       "nccl_bytes_received_per_rank_per_step": r["nccl_bytes_received_per_rank_per_step"],
                       "residual_Kalpha_minus_y_over_y": r["residual_Kalpha_minus_y_over_y"],
                       "phase_seconds_rank0_serialised": r.get("phase_seconds_rank0_serialised"), "parity": par},
            "clocks": r["clocks"], "gpu_launches": r["gpu_launches"],
            "e2e": {"value": flops / r["e2e_seconds_per_step"] * 1e-9, "unit": "GF/s", "seconds": r["e2e_seconds_per_step"],
                    "h2d_bytes_per_step": r["h2d_bytes_per_step"], "d2h_bytes_per_step": 2 * B.M_QUERY * 8,
                    "note": "every rank stages the replicated X, y, X* from pinned host memory and reads mean/variance back"},
            "roofline": B.step_roofline(n, sec, world, eng.ozaki)}
