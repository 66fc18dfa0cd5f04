"""CPU, world_size 2 over gloo: the host-side schedule of the block-row-cyclic sharded GP (ownership, packing, the
all-gather reorder, the reduce-based back-substitution).  The numerical block operations are INJECTED here by a
numpy/scipy checker (test infrastructure) -- the product uses CudaOps (battgp_b200/sharded.py) and is covered on GPUs by
tests/test_gpu_sharded.py."""
import math
import os
import socket

import numpy as np
import pytest
import scipy.linalg as sla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gp_oracle as orc


class CheckerOps:
    """Same interface as sharded.CudaOps, numpy on CPU tensors."""

    def __init__(self, spec):
        self.spec = spec

    def empty(self, rows, cols):
        return torch.zeros(rows, cols, dtype=torch.float64)

    def zeros_vec(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def scalars(self):
        return torch.full((1,), 2 ** 31 - 1, dtype=torch.int32), torch.zeros(1, dtype=torch.float64)

    def cov_block(self, xr, xc, out):
        out.copy_(torch.from_numpy(orc.cov(self.spec, xr.numpy(), xc.numpy())))

    def cov_diag(self, x):
        return torch.from_numpy(orc.cov_diag(self.spec, x.numpy()))

    def potrf_block(self, A, info, logdet):
        a = np.tril(A.numpy())
        L = sla.cholesky(a + np.tril(a, -1).T, lower=True)
        A.copy_(torch.from_numpy(L))
        logdet += 2 * np.log(np.diag(L)).sum()
        # same size as the engine's 128-block inverses; the checker's solves use L directly
        return torch.zeros(((A.shape[0] + 127) // 128) * 128 * 128, dtype=torch.float64)

    def trsm_rlt(self, L, dinv, X):
        X.copy_(torch.from_numpy(sla.solve_triangular(np.tril(L.numpy()), X.numpy().T, lower=True).T.copy()))

    def gemm_nt(self, A, B, C, alpha, beta, tri=False, roff=0, coff=0):
        r = alpha * (A.numpy() @ B.numpy().T) + beta * C.numpy()
        if tri:
            rr = np.arange(C.shape[0])[:, None] + roff
            cc = np.arange(C.shape[1])[None, :] + coff
            r = np.where(cc <= rr, r, C.numpy())
        C.copy_(torch.from_numpy(r))

    def trsv(self, L, dinv, b, trans):
        b.copy_(torch.from_numpy(sla.solve_triangular(np.tril(L.numpy()), b.numpy(), lower=True, trans="T" if trans else "N")))

    def gemv_t(self, A, v, y, alpha):
        y += alpha * torch.from_numpy(A.numpy().T @ v.numpy())

    def rowsumsq(self, V, out, accumulate):
        s = torch.from_numpy((V.numpy() ** 2).sum(1))
        if accumulate:
            out += s
        else:
            out.copy_(s)


def _worker(rank, world, port, n, nb, q, aug=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from battgp_b200 import engine as E
        from battgp_b200.sharded import ShardedGP
        x, y = orc.synth_field_data(n, seed=3)
        xq = orc.query_grid(x, 17)
        spec = orc.battgp_spec()
        gp = ShardedGP(E.battgp_spec(), torch.from_numpy(x), torch.from_numpy(y), 2.33e-6, nb=nb, ops=CheckerOps(spec))
        xq_t = torch.from_numpy(xq)
        gp.fit(xq_t if aug else None)            # aug: K(xq, X) and y ride through the factorisation as extra rows
        mean, var = gp.predict(xq_t)
        if aug:                                  # another grid after an augmented fit: the solve chain is still there
            m2, v2 = gp.predict(torch.from_numpy(xq[::2].copy()))
            assert torch.allclose(m2, mean[::2], rtol=1e-9, atol=0) and torch.allclose(v2, var[::2], rtol=1e-7, atol=0)
        owned_ok = all(i % world == rank for i in gp.owned) and sum(1 for _ in gp.owned) in (gp.nblk // world, gp.nblk // world + 1)
        assert gp.residual() < 1e-8
        q.put((rank, gp.lml, gp.alpha.numpy(), mean.numpy(), var.numpy(), owned_ok, gp.bytes_received))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n,nb,world,aug", [(700, 128, 2, False), (1000, 256, 2, True), (513, 128, 3, False), (513, 128, 3, True),
                                            (640, 128, 2, True)])
def test_sharded_schedule_matches_oracle_over_gloo(n, nb, world, aug):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, nb, q, aug)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y = orc.synth_field_data(n, seed=3)
    xq = orc.query_grid(x, 17)
    f = orc.fit(orc.battgp_spec(), x, y, 2.33e-6)
    mr, vr = orc.predict(orc.battgp_spec(), x, f, xq)
    for rank, lml, alpha, mean, var, owned_ok, recv in res:
        assert owned_ok
        assert abs(lml - f.lml) < 1e-9 * abs(f.lml)
        assert np.linalg.norm(alpha - f.alpha) / np.linalg.norm(f.alpha) < 1e-7
        np.testing.assert_allclose(mean, mr, rtol=1e-7)
        np.testing.assert_allclose(var, vr, rtol=1e-6)
        assert recv > 0


def test_single_process_schedule_without_process_group():
    from battgp_b200 import engine as E
    from battgp_b200.sharded import ShardedGP
    n = 300
    x, y = orc.synth_field_data(n, seed=1)
    gp = ShardedGP(E.battgp_spec(), torch.from_numpy(x), torch.from_numpy(y), 2.33e-6, nb=128, ops=CheckerOps(orc.battgp_spec()))
    gp.fit()
    f = orc.fit(orc.battgp_spec(), x, y, 2.33e-6)
    assert abs(gp.lml - f.lml) < 1e-9 * abs(f.lml)
    assert gp.owned == [0, 1, 2] and gp.bytes_received == 0
    # the same with the query rows riding through the factorisation
    xq = orc.query_grid(x, 11)
    gp2 = ShardedGP(E.battgp_spec(), torch.from_numpy(x), torch.from_numpy(y), 2.33e-6, nb=128, ops=CheckerOps(orc.battgp_spec()))
    gp2.fit(torch.from_numpy(xq))
    mean, var = gp2.predict(torch.from_numpy(xq))
    mr, vr = orc.predict(orc.battgp_spec(), x, f, xq)
    assert abs(gp2.lml - f.lml) < 1e-9 * abs(f.lml)
    np.testing.assert_allclose(mean.numpy(), mr, rtol=1e-7)
    np.testing.assert_allclose(var.numpy(), vr, rtol=1e-6)
