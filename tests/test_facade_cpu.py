"""CPU: host-side logic of the GPyTorch-shaped facade (no kernels are launched): constraint transforms, the
constraint-then-value set order of battcellgp_full.py:71-84, kernel-tree -> native spec flattening, gradient routing, the
import shim, and (when the reference tree is present, i.e. in the build container) that the UNMODIFIED reference modules
import and construct against it."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import battgp_b200.shim as shim

shim.install(force=True)
import gpytorch  # noqa: E402  (the stand-in)

from battgp_b200 import _lib  # noqa: E402
from battgp_b200.gpytorch.kernels import bind_spec  # noqa: E402

REF = "/root/reference"


def test_shim_registers_every_submodule_battgp_imports():
    import botorch
    from botorch.fit import fit_gpytorch_mll  # noqa: F401
    import gpytorch.constraints  # noqa: F401
    from gpytorch.utils.warnings import NumericalWarning  # noqa: F401
    assert gpytorch.__version__.endswith("battgp_b200")
    for name in ("models.ExactGP", "likelihoods.GaussianLikelihood", "means.ZeroMean", "means.Mean", "kernels.Kernel",
                 "kernels.RBFKernel", "kernels.ScaleKernel", "kernels.MultiDeviceKernel", "kernels.InducingPointKernel",
                 "kernels.MaternKernel", "kernels.PeriodicKernel", "distributions.MultivariateNormal",
                 "mlls.ExactMarginalLogLikelihood", "settings.fast_pred_var", "settings.debug", "constraints.Interval",
                 "constraints.Positive", "constraints.GreaterThan", "constraints.LessThan"):
        obj = gpytorch
        for part in name.split("."):
            obj = getattr(obj, part)
    assert hasattr(botorch.settings, "debug")


@pytest.mark.parametrize("cons,vals", [
    (lambda: gpytorch.constraints.Interval(1e-15, 1e4), [4.23e-13, 0.0099, 17.0]),
    (lambda: gpytorch.constraints.Interval(0.0, 1e5), [2.33e-6, 1.0]),
    (lambda: gpytorch.constraints.Positive(), [1e-3, 12.11, 400.0]),
    (lambda: gpytorch.constraints.GreaterThan(1e-4), [2e-4, 3.0]),
    (lambda: gpytorch.constraints.LessThan(5.0), [-3.0, 4.9]),
    (lambda: gpytorch.constraints.Interval(-math.inf, math.inf), [-2.0, 7.0]),
])
def test_constraint_roundtrip(cons, vals):
    c = cons()
    v = torch.tensor(vals, dtype=torch.float64)
    raw = c.inverse_transform(v)
    back = c.transform(raw)
    np.testing.assert_allclose(back.numpy(), v.numpy(), rtol=1e-10)
    assert c.check(v)


def test_interval_formula_is_sigmoid_scaled():
    c = gpytorch.constraints.Interval(2.0, 6.0)
    raw = torch.tensor([-1.0, 0.0, 2.0], dtype=torch.float64)
    np.testing.assert_allclose(c.transform(raw).numpy(), 2.0 + 4.0 / (1 + np.exp(-raw.numpy())), rtol=1e-14)
    p = gpytorch.constraints.Positive()
    np.testing.assert_allclose(p.transform(raw).numpy(), np.log1p(np.exp(raw.numpy())), rtol=1e-14)


def _kernel():
    class WienerKernel(gpytorch.kernels.Kernel):
        is_stationary = False

        def forward(self, x1, x2, **params):
            raise AssertionError("native: never evaluated in torch")
    kw = WienerKernel(active_dims=[0])
    kr = gpytorch.kernels.RBFKernel(ard_num_dims=3, active_dims=[1, 2, 3])
    return gpytorch.kernels.ScaleKernel(kw) + gpytorch.kernels.ScaleKernel(kr)


def test_bind_spec_flattens_the_battgp_kernel_tree():
    k = _kernel()
    k.kernels[0].outputscale = 4.0
    k.kernels[1].outputscale = 0.5
    k.kernels[1].base_kernel.lengthscale = torch.tensor([1.0, 2.0, 3.0])
    b = bind_spec(k, 4)
    spec = b.to_spec()
    assert [t.type for t in spec.terms] == [_lib.WIENER, _lib.RBF]
    assert [list(t.dims) for t in spec.terms] == [[0], [1, 2, 3]]
    assert abs(spec.terms[0].outputscale - 4.0) < 1e-6 and abs(spec.terms[1].outputscale - 0.5) < 1e-6
    np.testing.assert_allclose(spec.terms[1].lengthscale, [1.0, 2.0, 3.0], rtol=1e-6)
    # MultiDeviceKernel wrapper (cell_gp.py:38-43) is transparent
    md = gpytorch.kernels.MultiDeviceKernel(k, device_ids=range(2), output_device=torch.device("cpu"))
    assert [t.type for t in bind_spec(md, 4).to_spec().terms] == [_lib.WIENER, _lib.RBF]
    assert md.base_kernel.kernels[0].outputscale is not None
    # gradient routing: slots [os_w | os_r, ls0, ls1, ls2]
    g = torch.arange(1.0, 6.0, dtype=torch.float64)
    routed = b.route_grads(g)
    assert [tuple(r.shape) for r in routed] == [(), (), (1, 3)]
    assert float(routed[0]) == 1.0 and float(routed[1]) == 2.0 and routed[2].reshape(-1).tolist() == [3.0, 4.0, 5.0]


def test_bind_spec_isotropic_and_matern_periodic_and_unknown():
    iso = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel())
    iso.base_kernel.lengthscale = 2.0
    spec = bind_spec(iso, 3).to_spec()
    assert list(spec.terms[0].dims) == [0, 1, 2] and tuple(spec.terms[0].lengthscale) == pytest.approx((2.0, 2.0, 2.0))
    assert bind_spec(iso, 3).route_grads(torch.tensor([1.0, 1.0, 2.0, 3.0], dtype=torch.float64))[1].reshape(-1).tolist() == [6.0]
    mp = gpytorch.kernels.ScaleKernel(gpytorch.kernels.MaternKernel(nu=2.5, ard_num_dims=3, active_dims=[1, 2, 3])) + \
        gpytorch.kernels.ScaleKernel(gpytorch.kernels.PeriodicKernel(active_dims=[0]))
    assert [t.type for t in bind_spec(mp, 4).to_spec().terms] == [_lib.MATERN52, _lib.PERIODIC]

    class Odd(gpytorch.kernels.Kernel):
        def forward(self, x1, x2, **p):
            return x1 @ x2.T
    assert bind_spec(gpytorch.kernels.ScaleKernel(Odd()), 3) is None
    assert bind_spec(gpytorch.kernels.ScaleKernel(gpytorch.kernels.MaternKernel(nu=1.5)), 3) is None


def test_likelihood_constraint_then_value_order():
    lik = gpytorch.likelihoods.GaussianLikelihood()
    with pytest.raises(RuntimeError):
        lik.noise = 2.33e-6
    lik.noise_covar.raw_noise_constraint = gpytorch.constraints.Interval(0.0, 1e5)
    lik.noise = torch.tensor([2.33e-6])
    assert abs(float(lik.noise_covar) - 2.33e-6) < 1e-9


def test_exactgp_modes_and_lazy_prior_without_gpu():
    class M(gpytorch.models.ExactGP):
        def __init__(self, tx, ty):
            super().__init__(tx, ty, gpytorch.likelihoods.GaussianLikelihood())
            self.mean_module = gpytorch.means.ZeroMean()
            self.covar_module = _kernel()

        def forward(self, x):
            return gpytorch.distributions.MultivariateNormal(self.mean_module(x), self.covar_module(x))
    x = torch.randn(6, 4, dtype=torch.float64)
    m = M(x, torch.randn(6, dtype=torch.float64))
    assert isinstance(m.train_inputs, tuple) and m.train_inputs[0] is x
    m.train()
    out = m(x)                                        # prior: nothing evaluated, no GPU needed
    assert out.mean.shape == (6,) and tuple(out.lazy_covariance_matrix.shape) == (6, 6)
    with pytest.raises(RuntimeError):
        m(x[:3])
    with gpytorch.settings.debug(False):
        assert m(x[:3]).mean.shape == (3,)
    names = [n for n, _ in m.named_parameters()]
    assert names == ["likelihood.noise_covar.raw_noise", "covar_module.kernels.0.raw_outputscale",
                     "covar_module.kernels.1.raw_outputscale", "covar_module.kernels.1.base_kernel.raw_lengthscale"]
    m.double()
    assert m.train_targets.dtype == torch.float64
    if not torch.cuda.is_available():
        m.eval()
        with pytest.raises(_lib.BattGPLibraryError):
            m(x[:2])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_unmodified_reference_modules_construct_against_the_shim():
    sys.path.insert(0, REF)
    try:
        from src import config as cfg
        from src import gpytorch_utils
        from src.batt_models.battcellgp_full import BatteryCellGP_Full
        from src.gp import training  # noqa: F401   (imports botorch.fit)
        from src.gp.standard_models import ScaledRBFModel  # noqa: F401
        from src.gp.recursive_gp import RecursiveGP  # noqa: F401
        from src.gp.spatiotemporal_gp import ApproxSpatioTemporalGP  # noqa: F401
        assert isinstance(gpytorch_utils.get_scalar_gpytorch_constraint((0.0, math.inf)), gpytorch.constraints.Positive)
        assert isinstance(gpytorch_utils.get_scalar_gpytorch_constraint((1e-3, math.inf)), gpytorch.constraints.GreaterThan)
        assert isinstance(gpytorch_utils.get_vector_gpytorch_constraint([(1e-5, 1e4)] * 3), gpytorch.constraints.Interval)
        rng = np.random.default_rng(0)
        cell = BatteryCellGP_Full(rng.normal(size=(20, 4)), rng.normal(size=20), cellnr=1)      # CPU device, fp64
        p = cell.get_parameters()
        assert abs(float(cell.model.noise_variance.detach()) - cfg.NOISE_VARIANCE[0]) < 1e-9   # config.py:39 is a 1-tuple
        assert abs(float(cell.model.outputscale_rbf.detach()) - cfg.OUTPUTSCALE_RBF) < 1e-7
        np.testing.assert_allclose(cell.model.lengthscale_rbf.detach().numpy().reshape(-1), cfg.LENGTHSCALE_RBF, rtol=1e-6)
        spec = bind_spec(cell.model.covar_module, 4).to_spec()
        assert [t.type for t in spec.terms] == [_lib.WIENER, _lib.RBF] and p["n_devices"] == 1
        xt, yt = cell.get_training_data()
        assert xt.shape == (20, 4) and yt.shape == (20,)
        with pytest.raises(ValueError):
            BatteryCellGP_Full(xt, yt, bogus=1)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
