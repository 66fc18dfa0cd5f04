"""GPU: the GPyTorch-shaped facade (battgp_b200.gpytorch / .botorch) used the way /root/reference/src uses GPyTorch --
model classes below restate the *API usage pattern* of cell_gp.py / standard_models.py / training.py (the reference tree
is not on the GPU box) -- checked against the oracle and the golden vectors produced by the reference's own code."""
import math
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import gp_oracle as orc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def gpytorch():
    import battgp_b200.shim as shim
    shim.install(force=True)
    import gpytorch
    return gpytorch


def _wiener_cls(gpytorch, name="WienerKernel"):
    """A user-defined integrated-Wiener kernel written against the public Kernel API (vectorised; same formula as
    wiener_kernel.py:32).  Named ``WienerKernel`` it is recognised as a native term; under another name it exercises
    the generic path (user torch forward + engine factorisation)."""
    def forward(self, x1, x2, **params):
        dist = self.covar_dist(x1, x2, **params)
        m = torch.minimum(x1, x2.reshape(1, -1)) if not params.get("diag", False) else torch.minimum(x1, x2)
        return m.pow(3) / 3 + dist * m.pow(2) / 2
    return type(name, (gpytorch.kernels.Kernel,), {"is_stationary": False, "forward": forward})


def _cell_model(gpytorch, x, y, theta, wiener_name="WienerKernel"):
    """Same construction sequence as cell_gp.py:13-49 + battcellgp_full.py:71-84 (constraints first, then values)."""
    W = _wiener_cls(gpytorch, wiener_name)

    class CellGP(gpytorch.models.ExactGP):
        def __init__(self, tx, ty):
            super().__init__(tx, ty, likelihood=gpytorch.likelihoods.GaussianLikelihood())
            self.mean_module = gpytorch.means.ZeroMean()
            kw = W(active_dims=[0])
            kr = gpytorch.kernels.RBFKernel(ard_num_dims=3, active_dims=[1, 2, 3])
            self.covar_module = gpytorch.kernels.ScaleKernel(kw) + gpytorch.kernels.ScaleKernel(kr)
            self.to(tx.device)

        def forward(self, xx):
            return gpytorch.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    m = CellGP(x, y)
    dev = x.device
    m.likelihood.noise_covar.raw_noise_constraint = gpytorch.constraints.Interval(0.0, 1e5).to(dev)
    m.likelihood.noise = torch.tensor([theta["noise"]]).to(dev)
    m.covar_module.kernels[0].raw_outputscale_constraint = gpytorch.constraints.Interval(1e-15, 1e4).to(dev)
    m.covar_module.kernels[0].outputscale = torch.tensor([theta["os_w"]]).to(dev)
    m.covar_module.kernels[1].raw_outputscale_constraint = gpytorch.constraints.Interval(1e-15, 1e4).to(dev)
    m.covar_module.kernels[1].outputscale = torch.tensor([theta["os_r"]]).to(dev)
    m.covar_module.kernels[1].base_kernel.lengthscale = torch.tensor(theta["ls"]).to(dev)
    m.eval(); m.likelihood.eval()
    return m


THETA = {"noise": 2.33e-6, "os_w": 4.23e-13, "os_r": 0.0099, "ls": [12.11, 33.75, 45.14]}


@pytest.fixture()
def f64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)      # gp_runner.py:158-159 does this in its workers
    yield
    torch.set_default_dtype(old)


def test_scaled_rbf_known_answers_fp32(gpytorch):
    """test_standard_models.py:12-47 restated: fp32 model, one and two training points."""
    class ScaledRBF(gpytorch.models.ExactGP):
        def __init__(self, tx, ty, nv, s, l):
            lik = gpytorch.likelihoods.GaussianLikelihood()
            super().__init__(tx.type(torch.float32), ty.type(torch.float32), lik)
            self.mean_module = gpytorch.means.ZeroMean()
            self.covar_module = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel())
            self.likelihood.noise = nv
            self.covar_module.outputscale = s
            self.covar_module.base_kernel.lengthscale = l
            self.eval()

        def forward(self, xx):
            return gpytorch.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    x = torch.Tensor([[1.0]])
    gp = ScaledRBF(x, torch.Tensor([10.0]), 3.0, 3.0, 2.0)
    with pytest.warns(gpytorch.utils.warnings.GPInputWarning):
        gp(x)
    with torch.no_grad(), gpytorch.settings.fast_pred_var(), gpytorch.settings.debug(False), warnings.catch_warnings():
        warnings.simplefilter("error")
        out = gp(x)
    assert out.mean.dtype == torch.float32 and out.mean.device.type == "cpu"     # CPU tensors in, CPU tensors out
    assert abs(float(out.mean[0]) - 5.0) < 1e-6
    assert abs(float(np.diag(out._covar.detach().cpu().numpy())[0]) - 1.5) < 1e-5
    gp = ScaledRBF(torch.Tensor([[1.0], [1.0]]), torch.Tensor([10.0, 10.0]), 3.0, 3.0, 2.0)
    with torch.no_grad(), gpytorch.settings.debug(False):
        out = gp(x)
    assert abs(float(out.mean[0]) - (5.0 / 1.5 + 10.0 / 3.0) / (1 / 1.5 + 1 / 3.0)) < 1e-5
    assert abs(float(out.variance[0]) - 1.0) < 1e-5


def test_full_covariance_matches_reference_recursive_gp(gpytorch, f64_default):
    """test_recursive_gp.py:195-232: exact GP == recursive GP (golden from the reference's recursive_gp.py), rel 1e-5."""
    class M(gpytorch.models.ExactGP):
        def __init__(self, tx, ty):
            super().__init__(tx, ty, gpytorch.likelihoods.GaussianLikelihood())
            self.mean_module = gpytorch.means.ZeroMean()
            self.covar_module = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel())
            self.likelihood.noise = 3.0
            self.covar_module.outputscale = 3.0
            self.covar_module.base_kernel.lengthscale = 2.0
            self.eval()

        def forward(self, xx):
            return gpytorch.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    gp = M(torch.tensor(G["rgp_rbf_xt"]), torch.tensor(G["rgp_rbf_yt"]))
    with torch.no_grad():
        out = gp(torch.tensor(G["rgp_rbf_xq"]))
    yq = out.mean.numpy()
    cq = out._covar.detach().cpu().numpy()
    assert np.linalg.norm(yq - G["rgp_rbf_mean"]) < 1e-5 * np.linalg.norm(G["rgp_rbf_mean"])
    assert np.linalg.norm(cq - G["rgp_rbf_cov"]) < 1e-5 * np.linalg.norm(G["rgp_rbf_cov"])


@pytest.mark.parametrize("wname", ["WienerKernel", "MyIntegratedWiener"])
def test_wiener_rbf_exact_gp_equals_reference_kalman_stgp(gpytorch, f64_default, wname):
    """test_spatiotemporal_gp.py:218-282 restated with the golden STGP outputs (rel 1e-6), native and generic Wiener."""
    tt, st, yt, sq = G["stgp_t"], G["stgp_s"], G["stgp_y"], G["stgp_sq"]
    xt = np.hstack([tt[:, None], st])
    th = {"noise": 0.1, "os_w": 10.0, "os_r": 3.0, "ls": [2.0, 2.0, 2.0]}
    for i in (0, 3, len(tt) - 1):
        m = _cell_model(gpytorch, torch.tensor(xt[: i + 1]), torch.tensor(yt[: i + 1]), th, wname)
        xq = np.hstack([np.full((sq.shape[0], 1), tt[i]), sq])
        with torch.no_grad(), gpytorch.settings.fast_pred_var():
            out = m(torch.tensor(xq))
        assert np.linalg.norm(out.mean.numpy() - G["stgp_mean"][i]) < 1e-6 * np.linalg.norm(G["stgp_mean"][i])
        assert np.linalg.norm(out.variance.numpy() - G["stgp_var"][i]) < 1e-6 * np.linalg.norm(G["stgp_var"][i])


def test_cell_model_predict_matches_oracle_on_device(gpytorch, f64_default):
    """battcellgp_full.py:168-195 flow: tensors on the GPU, default hyper-parameters, 300 query points."""
    x, y = orc.synth_field_data(3000, seed=0)
    xq = orc.query_grid(x)
    m = _cell_model(gpytorch, torch.tensor(x, device=DEV), torch.tensor(y, device=DEV), THETA)
    with torch.no_grad(), gpytorch.settings.fast_pred_var():
        out = m(torch.tensor(xq, device=DEV).contiguous())
    f = orc.fit(orc.battgp_spec(), x, y, 2.33e-6)
    mr, vr = orc.predict(orc.battgp_spec(), x, f, xq)
    np.testing.assert_allclose(out.mean.detach().cpu().numpy(), mr, rtol=1e-7)
    np.testing.assert_allclose(out.variance.detach().cpu().numpy().reshape(-1), vr, rtol=1e-6)
    # second call reuses the cached factor; changing a hyper-parameter invalidates it
    st0 = m.prediction_strategy[1]
    with torch.no_grad():
        m(torch.tensor(xq[:5], device=DEV))
    assert m.prediction_strategy[1] is st0
    m.covar_module.kernels[1].outputscale = torch.tensor([0.02], device=DEV)
    with torch.no_grad():
        out2 = m(torch.tensor(xq, device=DEV))
    assert m.prediction_strategy[1] is not st0
    f2 = orc.fit(orc.battgp_spec(outputscale_rbf=0.02), x, y, 2.33e-6)
    m2, _ = orc.predict(orc.battgp_spec(outputscale_rbf=0.02), x, f2, xq)
    np.testing.assert_allclose(out2.mean.cpu().numpy(), m2, rtol=1e-7)


def test_kernel_forward_dense_cross_covariance(gpytorch, f64_default):
    """recursive_gp.py:52-57 / spatiotemporal_gp.py:157-162 call kernel.forward(x1, x2[, diag]) on CPU tensors."""
    x1, _ = orc.synth_field_data(70, seed=1)
    x2, _ = orc.synth_field_data(33, seed=2)
    m = _cell_model(gpytorch, torch.tensor(x1), torch.zeros(70), THETA)
    k = m.covar_module.forward(torch.tensor(x1), torch.tensor(x2))
    assert torch.is_tensor(k) and k.device.type == "cpu"
    np.testing.assert_allclose(k.numpy(), orc.cov(orc.battgp_spec(), x1, x2), rtol=1e-12, atol=1e-20)
    d = m.covar_module.forward(torch.tensor(x1), torch.tensor(x1), diag=True)
    np.testing.assert_allclose(d.detach().numpy(), orc.cov_diag(orc.battgp_spec(), x1), rtol=1e-12)
    kr = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=3))
    kr.outputscale = torch.tensor(3.0); kr.base_kernel.lengthscale = torch.tensor(2.0)
    kk = kr.forward(torch.tensor(x1[:, 1:]), torch.tensor(x2[:, 1:]))
    np.testing.assert_allclose(kk.detach().numpy(), orc.cov(orc.scaled_rbf_spec(3, 3.0, 2.0), x1[:, 1:], x2[:, 1:]), rtol=1e-12)


def _expected_raw_grads(model, gpytorch, x, y, theta_names):
    """-(1/N) dLML/dtheta (oracle) chained through the constraint transforms with torch autograd."""
    n = x.shape[0]
    lk = model.likelihood.noise_covar
    k0, k1 = model.covar_module.kernels[0], model.covar_module.kernels[1]
    vals = dict(noise=float(lk.noise), os_w=float(k0.outputscale), os_r=float(k1.outputscale),
                ls=k1.base_kernel.lengthscale.detach().cpu().reshape(-1).tolist())
    g = orc.lml_grad(orc.battgp_spec(vals["os_w"], vals["os_r"], tuple(vals["ls"])), x, y, vals["noise"])
    exp = {}
    for name, raw, cons, dth in (("noise", lk.raw_noise, lk.raw_noise_constraint, [g["noise"]]),
                                 ("os_w", k0.raw_outputscale, k0.raw_outputscale_constraint, [g["terms"][0]["outputscale"]]),
                                 ("os_r", k1.raw_outputscale, k1.raw_outputscale_constraint, [g["terms"][1]["outputscale"]]),
                                 ("ls", k1.base_kernel.raw_lengthscale, k1.base_kernel.raw_lengthscale_constraint, g["terms"][1]["lengthscale"])):
        r = raw.detach().clone().cpu().double().requires_grad_(True)
        th = cons.cpu().transform(r) if hasattr(cons, "cpu") else cons.transform(r)
        th.backward(torch.tensor(dth, dtype=torch.float64).reshape(th.shape))
        exp[name] = (-r.grad / n).reshape(-1).numpy()
        cons.to(raw.device)
    return exp, -g["lml"] / n


def test_mll_value_and_backward_match_oracle(gpytorch, f64_default):
    """training.py:39-41: output = model(train_x); loss = -mll(output, y); loss.backward()."""
    x, y = orc.synth_field_data(700, seed=6)
    xt, yt = torch.tensor(x, device=DEV), torch.tensor(y, device=DEV)
    th = dict(THETA, ls=[10.0, 30.0, 50.0], noise=5e-6)
    m = _cell_model(gpytorch, xt, yt, th)
    m.train(); m.likelihood.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m)
    loss = -mll(m(xt), yt)
    loss.backward()
    exp, loss_ref = _expected_raw_grads(m, gpytorch, x, y, None)
    assert abs(float(loss) - loss_ref) < 1e-9 * abs(loss_ref)
    got = {"noise": m.likelihood.noise_covar.raw_noise.grad, "os_w": m.covar_module.kernels[0].raw_outputscale.grad,
           "os_r": m.covar_module.kernels[1].raw_outputscale.grad, "ls": m.covar_module.kernels[1].base_kernel.raw_lengthscale.grad}
    for k, e in exp.items():
        gk = got[k].detach().cpu().reshape(-1).numpy()
        scale = max(np.abs(e).max(), 1e-8)
        assert np.max(np.abs(gk - e)) < 1e-5 * scale + 1e-12, (k, gk, e)
    with pytest.raises(RuntimeError):
        m(xt[:10])                                    # "You must train on the training inputs!"


def test_generic_kernel_mll_backward_matches_native(gpytorch, f64_default):
    x, y = orc.synth_field_data(300, seed=8)
    xt, yt = torch.tensor(x, device=DEV), torch.tensor(y, device=DEV)
    grads = {}
    for name in ("WienerKernel", "OtherWiener"):
        m = _cell_model(gpytorch, xt, yt, dict(THETA, noise=5e-6), name)
        m.train(); m.likelihood.train()
        mll = gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m)
        loss = -mll(m(xt), yt)
        loss.backward()
        grads[name] = (float(loss), torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy())
    assert abs(grads["WienerKernel"][0] - grads["OtherWiener"][0]) < 1e-9 * abs(grads["OtherWiener"][0])
    a, b = grads["WienerKernel"][1], grads["OtherWiener"][1]
    assert np.max(np.abs(a - b)) < 1e-5 * np.abs(b).max()


def test_training_loops_decrease_the_loss(gpytorch, f64_default):
    """training.py:11-67 (Adam), :108-171 (L-BFGS strong Wolfe), :70-105 (botorch fit_gpytorch_mll)."""
    import botorch
    x, y = orc.synth_field_data(400, seed=12)
    xt, yt = torch.tensor(x, device=DEV), torch.tensor(y, device=DEV)
    start = dict(THETA, ls=[5.0, 10.0, 15.0], os_r=0.05, noise=1e-5)

    def loss_of(m):
        m.train(); m.likelihood.train()
        mll = gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m)
        with torch.no_grad():
            return float(-mll(m(xt), yt))

    m = _cell_model(gpytorch, xt, yt, start)
    l0 = loss_of(m)
    opt = torch.optim.Adam(m.parameters(), lr=0.1)
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m)
    for _ in range(15):
        opt.zero_grad()
        loss = -mll(m(xt), yt)
        loss.backward()
        opt.step()
    assert loss_of(m) < l0

    m = _cell_model(gpytorch, xt, yt, start)
    m.train(); m.likelihood.train()
    opt = torch.optim.LBFGS(m.parameters(), line_search_fn="strong_wolfe", lr=1)
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m)

    def closure():
        opt.zero_grad()
        loss = -mll(m(xt), yt)
        loss.backward()
        return loss
    for _ in range(3):
        opt.step(closure)
    l_lbfgs = loss_of(m)
    assert l_lbfgs < l0

    m = _cell_model(gpytorch, xt, yt, start)
    m.train(); m.likelihood.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(m.likelihood, m).to(xt)
    botorch.fit.fit_gpytorch_mll(mll, max_retries=1, optimizer_kwargs={"options": {"maxiter": 30, "ftol": 1e-15, "gtol": 1e-15}})
    assert not m.training
    assert loss_of(m) < l0


def test_constraint_errors_and_set_order(gpytorch):
    lik = gpytorch.likelihoods.GaussianLikelihood()
    with pytest.raises(RuntimeError):
        lik.noise = 2.33e-6                 # below the default GreaterThan(1e-4): battcellgp_full.py:71 swaps it first
    lik.noise_covar.raw_noise_constraint = gpytorch.constraints.Interval(0.0, 1e5)
    lik.noise = 2.33e-6
    assert abs(float(lik.noise) - 2.33e-6) < 1e-9


def test_inducing_point_kernel_sgpr_prediction(gpytorch):
    """standard_models.py:58-107 (SparseScaledRBFModel): ExactGP over InducingPointKernel = subset-of-regressors prediction,
    fp32 tensors in and out, checked against the closed form evaluated with numpy in fp64."""
    rng = np.random.default_rng(11)
    n, m, q = 300, 30, 25
    x = rng.normal(size=(n, 2)).astype(np.float32); y = (np.sin(2 * x[:, 0]) + 0.1 * rng.normal(size=n)).astype(np.float32)
    u = x[rng.choice(n, m, replace=False)]; xq = rng.normal(size=(q, 2)).astype(np.float32)

    class Sparse(gpytorch.models.ExactGP):
        def __init__(self, tx, ty, ind):
            lik = gpytorch.likelihoods.GaussianLikelihood()
            super().__init__(tx, ty, lik)
            self.mean_module = gpytorch.means.ZeroMean()
            self.base_covar_module = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel())
            self.likelihood.noise = 0.05
            self.base_covar_module.outputscale = 1.3
            self.base_covar_module.base_kernel.lengthscale = 0.9
            self.covar_module = gpytorch.kernels.InducingPointKernel(self.base_covar_module, inducing_points=ind, likelihood=lik)
            self.eval()

        def forward(self, xx):
            return gpytorch.distributions.MultivariateNormal(self.mean_module(xx), self.covar_module(xx))

    gp = Sparse(torch.tensor(x), torch.tensor(y), torch.tensor(u))
    with torch.no_grad(), gpytorch.settings.fast_pred_var(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = gp(torch.tensor(xq))
    mean = out.mean.numpy().astype(np.float64)
    var = np.diag(out._covar.detach().cpu().numpy()).astype(np.float64)
    assert out.mean.dtype == torch.float32
    spec = orc.scaled_rbf_spec(2, float(np.float32(1.3)), float(np.float32(0.9)))
    x64, u64, q64, y64 = (a.astype(np.float64) for a in (x, u, xq, y))
    Kuu = orc.cov(spec, u64, u64); Kfu = orc.cov(spec, x64, u64); Ksu = orc.cov(spec, q64, u64)
    W = np.linalg.solve(Kuu, Kfu.T)
    G_ = Kfu @ W + float(np.float32(0.05)) * np.eye(n)
    Qsf = Ksu @ W
    mref = Qsf @ np.linalg.solve(G_, y64)
    vref = np.diag(Ksu @ np.linalg.solve(Kuu, Ksu.T) - Qsf @ np.linalg.solve(G_, Qsf.T))
    np.testing.assert_allclose(mean, mref, atol=2e-4)
    np.testing.assert_allclose(var, vref, atol=2e-4)
