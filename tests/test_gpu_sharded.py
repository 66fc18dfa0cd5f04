"""GPU: the sharded GP's block operations through the C-ABI (CudaOps).  With one visible GPU the schedule runs with
P = 1 (every stripe local, no collective); with >= 2 GPUs a 2-rank NCCL run is compared against the single-GPU engine."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu


def _t(a, dev="cuda:0"):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)


@pytest.mark.parametrize("n,nb,aug", [(1000, 256, False), (3000, 512, False), (2500, 1024, False), (3000, 512, True), (2500, 1024, True),
                                      (4100, 2048, True)])
def test_sharded_single_rank_matches_oracle_and_engine(eng, n, nb, aug):
    """aug: K(xq, X) and y ride through the factorisation as extra rows (int8 products for the 64-aligned panels)."""
    from battgp_b200 import engine as E
    from battgp_b200.sharded import ShardedGP
    x, y = orc.synth_field_data(n, seed=5)
    xq = orc.query_grid(x)
    xq_t = _t(xq)
    gp = ShardedGP(E.battgp_spec(), _t(x), _t(y), 2.33e-6, nb=nb).fit(xq_t if aug else None)
    mean, var = gp.predict(xq_t)
    f = orc.fit(orc.battgp_spec(), x, y, 2.33e-6)
    mr, vr = orc.predict(orc.battgp_spec(), x, f, xq)
    assert abs(gp.lml - f.lml) < 1e-9 * abs(f.lml)
    assert np.linalg.norm(gp.alpha.cpu().numpy() - f.alpha) / np.linalg.norm(f.alpha) < 1e-7
    np.testing.assert_allclose(mean.cpu().numpy(), mr, rtol=1e-7)
    np.testing.assert_allclose(var.cpu().numpy(), vr, rtol=1e-6)
    assert gp.residual() < 1e-8
    # the stripes hold the same factor as the monolithic engine
    st = E.fit(E.battgp_spec(), _t(x), _t(y), 2.33e-6)
    L = torch.tril(st.L)
    for i in gp.owned:
        b0, e = gp.b0(i), gp.e(i)
        blk = gp.rows[i].clone()
        ref = L[b0:e, :e]
        blk[:, b0:e] = torch.tril(blk[:, b0:e])
        assert ((blk - ref).norm() / ref.norm()).item() < 1e-10


def test_trsv_gemv_t_rowsumsq_blocks(eng):
    rng = np.random.default_rng(0)
    n = 700
    a = rng.normal(size=(n, n)); k = a @ a.T + n * np.eye(n)
    A = _t(k)
    info, _, dinv = eng.potrf(A)
    assert info == 0
    L = np.tril(A.cpu().numpy())
    b = rng.normal(size=n)
    import scipy.linalg as sla
    x1 = eng.trsv(A, dinv, _t(b), False).cpu().numpy()
    x2 = eng.trsv(A, dinv, _t(b), True).cpu().numpy()
    assert np.linalg.norm(x1 - sla.solve_triangular(L, b, lower=True)) / np.linalg.norm(x1) < 1e-10
    assert np.linalg.norm(x2 - sla.solve_triangular(L, b, lower=True, trans="T")) / np.linalg.norm(x2) < 1e-10
    M = rng.normal(size=(333, 517)); v = rng.normal(size=333); yv = rng.normal(size=517)
    out = eng.gemv_t(_t(M), _t(v), _t(yv), -0.5).cpu().numpy()
    np.testing.assert_allclose(out, yv - 0.5 * M.T @ v, rtol=1e-12, atol=1e-12)
    acc = _t(np.ones(333))
    eng.rowsumsq(_t(M), acc, True)
    np.testing.assert_allclose(acc.cpu().numpy(), 1 + (M ** 2).sum(1), rtol=1e-13)


def _collect(q, ps, n, timeout=300.0):
    """n results from the workers' queue; fails as soon as a worker has died instead of waiting out the timeout."""
    import queue as _queue
    import time as _time
    res, t0 = [], _time.monotonic()
    while len(res) < n:
        try:
            res.append(q.get(timeout=2.0))
        except _queue.Empty:
            dead = [p.exitcode for p in ps if p.exitcode not in (None, 0)]
            if dead:
                for p in ps:
                    if p.is_alive():
                        p.terminate()
                pytest.fail(f"worker process exited with {dead} before delivering its result")
            if _time.monotonic() - t0 > timeout:
                for p in ps:
                    if p.is_alive():
                        p.terminate()
                pytest.fail("timed out waiting for the worker processes")
    return res



def _nccl_worker(rank, world, port, n, nb, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from battgp_b200 import engine as E
        from battgp_b200.sharded import ShardedGP
        x, y = orc.synth_field_data(n, seed=5)
        dev = f"cuda:{rank}"
        xq_t = _t(orc.query_grid(x), dev)
        gp = ShardedGP(E.battgp_spec(), _t(x, dev), _t(y, dev), 2.33e-6, nb=nb).fit(xq_t)      # query rows + y ride along
        mean, var = gp.predict(xq_t)
        m2, v2 = gp.predict(xq_t[::3].contiguous())                                               # another grid: solve chain
        assert torch.allclose(m2, mean[::3], rtol=1e-9, atol=0) and torch.allclose(v2, var[::3], rtol=1e-6, atol=0)
        q.put((rank, gp.lml, mean.cpu().numpy(), var.cpu().numpy(), gp.bytes_received))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_two_ranks_nccl_matches_oracle():
    import torch.multiprocessing as mp
    n, nb, world = 5000, 512, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_nccl_worker, args=(r, world, port, n, nb, q)) for r in range(world)]
    [p.start() for p in ps]
    res = _collect(q, ps, world)
    [p.join(timeout=60) for p in ps]
    x, y = orc.synth_field_data(n, seed=5)
    f = orc.fit(orc.battgp_spec(), x, y, 2.33e-6)
    mr, vr = orc.predict(orc.battgp_spec(), x, f, orc.query_grid(x))
    for rank, lml, mean, var, recv in res:
        assert abs(lml - f.lml) < 1e-9 * abs(f.lml)
        np.testing.assert_allclose(mean, mr, rtol=1e-7)
        np.testing.assert_allclose(var, vr, rtol=1e-6)
        assert recv > 0


def _facade_worker(rank, world, port, q):
    """Every rank builds the SAME model with MultiDeviceKernel(device_ids=range(world)) -- the reference's only intra-GP
    multi-device knob (cell_gp.py:38-43) -- and calls it in eval mode; the facade routes it to the sharded engine."""
    import sys
    import warnings
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, os.path.join(root, "tools"))
        import run_reference_modules as rrm
        torch.set_default_dtype(torch.float64)
        g = np.load(os.path.join(root, "tests", "golden", "real_field_data.npz"))
        x, y, xq = g["b14_c3_x"], g["b14_c3_y"], g["b14_c3_xq"]
        dev = torch.device("cuda", rank)
        ref = rrm.find_reference()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if ref is not None:                      # the reference's own classes, unmodified
                rrm.install(ref)
                from src.batt_models.battcellgp_full import BatteryCellGP_Full
                cell = BatteryCellGP_Full(x, y, 3, device=dev, n_devices=world, output_device=dev)
                mean, var = cell.predict(xq)
                used = "reference BatteryCellGP_Full(n_devices=%d)" % world
                strat = cell.model.prediction_strategy
            else:
                import battgp_b200.shim as shim
                shim.install(force=True)
                import gpytorch

                class M(gpytorch.models.ExactGP):
                    def __init__(self, tx, ty):
                        super().__init__(tx, ty, gpytorch.likelihoods.GaussianLikelihood(noise_constraint=gpytorch.constraints.Interval(0.0, 1e5)))
                        self.mean_module = gpytorch.means.ZeroMean()
                        k = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=3, active_dims=[1, 2, 3]))
                        self.covar_module = gpytorch.kernels.MultiDeviceKernel(k, device_ids=range(world), output_device=dev)
                        self.to(tx.device)

                    def forward(self, xx):
                        return gpytorch.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))
                m = M(_t(x, dev), _t(y, dev))
                m.likelihood.noise = torch.tensor([2.33e-6], device=dev)
                m.covar_module.base_kernel.outputscale = torch.tensor(0.0099, device=dev)
                m.covar_module.base_kernel.base_kernel.lengthscale = torch.tensor([12.11, 33.75, 45.14], device=dev)
                m.eval()
                out = m(_t(xq, dev))
                mean, var = out.mean.cpu().numpy(), out.variance.cpu().numpy()
                used = "local ExactGP with MultiDeviceKernel"
                strat = m.prediction_strategy
        q.put((rank, used, np.asarray(mean), np.asarray(var), type(strat[1]).__name__))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multi_device_kernel_routes_to_the_sharded_engine():
    import torch.multiprocessing as mp
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_facade_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = _collect(q, ps, world)
    [p.join(timeout=60) for p in ps]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "real_field_data.npz"))
    for rank, used, mean, var, strat in res:
        assert strat == "ShardedGP", (used, strat)
        if used.startswith("reference"):
            np.testing.assert_allclose(mean, g["b14_c3_mean"], rtol=1e-6)
            np.testing.assert_allclose(var, g["b14_c3_var"], rtol=1e-4)
        else:
            spec = orc.scaled_rbf_spec(3, 0.0099, 1.0)
            spec.terms[0].dims = [1, 2, 3]; spec.terms[0].lengthscale = (12.11, 33.75, 45.14)
            f = orc.fit(spec, g["b14_c3_x"], g["b14_c3_y"], 2.33e-6)
            mr, vr = orc.predict(spec, g["b14_c3_x"], f, g["b14_c3_xq"])
            np.testing.assert_allclose(mean, mr, rtol=1e-6)
            np.testing.assert_allclose(var, vr, rtol=1e-4)
