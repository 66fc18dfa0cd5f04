"""CPU: pin the oracle (oracle/gp_oracle.py) against (a) golden vectors produced by the reference's own code
(tests/golden/make_golden.py), (b) the reference's analytic known-answer tests, (c) self-consistency (finite differences,
identities) for the quantities no reference test pins (LML, gradient, L, alpha)."""
import math
import os

import numpy as np
import pytest

from oracle import gp_oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


def test_wiener_cov_matches_reference_forward():
    # /root/reference/src/gp/wiener_kernel.py:10-32 executed verbatim by make_golden.py
    np.testing.assert_allclose(orc.wiener_cov(G["wiener_t1"], G["wiener_t2"]), G["wiener_cross"], rtol=1e-14, atol=0)
    np.testing.assert_allclose(orc.wiener_cov(G["wiener_t1"], G["wiener_t1"]), G["wiener_self"], rtol=1e-14, atol=0)
    spec = orc.KernelSpec([orc.Term(orc.WIENER, [0], 1.0)])
    np.testing.assert_allclose(orc.cov_diag(spec, G["wiener_t1"]), G["wiener_diag"], rtol=1e-14)


def test_scaled_rbf_analytic_known_answers():
    # /root/reference/tests/gp/test_standard_models.py:12-47 (reference runs them in fp32 to 5 decimals)
    spec = orc.scaled_rbf_spec(1, 3.0, 2.0)
    x = np.array([[1.0]]); y = np.array([10.0])
    f = orc.fit(spec, x, y, 3.0)
    m, v = orc.predict(spec, x, f, x, clamp=False)
    assert abs(m[0] - 5.0) < 1e-12 and abs(v[0] - 1.5) < 1e-12
    x = np.array([[1.0], [1.0]]); y = np.array([10.0, 10.0])
    f = orc.fit(spec, x, y, 3.0)
    m, v = orc.predict(spec, x, f, np.array([[1.0]]), clamp=False)
    # second observation at the same location: the reference's expected values (test_standard_models.py:46-47)
    assert abs(m[0] - (5 / 1.5 + 10 / 3) / (1 / 1.5 + 1 / 3)) < 1e-12 and abs(v[0] - 1.0) < 1e-12


def test_recursive_gp_analytic_rbf_values():
    # /root/reference/tests/gp/test_recursive_gp.py:48-193 pins kernel.forward values, e.g. var = 3 - 3 exp(-d^2/4)
    spec = orc.scaled_rbf_spec(1, 3.0, 2.0)
    d = 1.7
    k = orc.cov(spec, np.array([[0.0]]), np.array([[d]]))[0, 0]
    assert abs(k - 3 * math.exp(-d * d / 8)) < 1e-15
    f = orc.fit(spec, np.array([[0.0]]), np.array([1.0]), 0.0 + 1e-300)
    _, v = orc.predict(spec, np.array([[0.0]]), f, np.array([[d]]), clamp=False)
    assert abs(v[0] - (3 - 3 * math.exp(-d * d / 4))) < 1e-12


def test_exact_gp_equals_reference_recursive_gp_rbf_full_cov():
    # golden from /root/reference/src/gp/recursive_gp.py; tolerance of the reference's own test (rel 1e-5)
    spec = orc.scaled_rbf_spec(3, 3.0, 2.0)
    f = orc.fit(spec, G["rgp_rbf_xt"], G["rgp_rbf_yt"], 3.0)
    m, c = orc.predict(spec, G["rgp_rbf_xt"], f, G["rgp_rbf_xq"], full_cov=True)
    assert np.linalg.norm(m - G["rgp_rbf_mean"]) < 1e-5 * np.linalg.norm(G["rgp_rbf_mean"])
    assert np.linalg.norm(c - G["rgp_rbf_cov"]) < 1e-5 * np.linalg.norm(G["rgp_rbf_cov"])


def test_exact_gp_equals_reference_recursive_gp_wiener_rbf():
    th = G["rgp_wr_theta"]
    spec = orc.battgp_spec(th[0], th[1], tuple(th[2:5]))
    f = orc.fit(spec, G["rgp_wr_xt"], G["rgp_wr_yt"], th[5])
    m, v = orc.predict(spec, G["rgp_wr_xt"], f, G["rgp_wr_xq"], clamp=False)
    assert np.linalg.norm(m - G["rgp_wr_mean"]) < 1e-5 * np.linalg.norm(G["rgp_wr_mean"])
    assert np.linalg.norm(v - G["rgp_wr_var"]) < 1e-5 * np.linalg.norm(G["rgp_wr_var"])


def test_exact_gp_equals_reference_kalman_stgp_every_step():
    # golden from /root/reference/src/gp/spatiotemporal_gp.py; tolerance of test_spatiotemporal_gp.py:218-282 (1e-6)
    spec = orc.battgp_spec(10.0, 3.0, (2.0, 2.0, 2.0))
    tt, st, yt, sq = G["stgp_t"], G["stgp_s"], G["stgp_y"], G["stgp_sq"]
    xt = np.hstack([tt[:, None], st])
    for i in range(len(tt)):
        f = orc.fit(spec, xt[: i + 1], yt[: i + 1], 0.1)
        xq = np.hstack([np.full((sq.shape[0], 1), tt[i]), sq])
        m, v = orc.predict(spec, xt[: i + 1], f, xq, clamp=False)
        assert np.linalg.norm(m - G["stgp_mean"][i]) < 1e-6 * np.linalg.norm(G["stgp_mean"][i]), i
        assert np.linalg.norm(v - G["stgp_var"][i]) < 1e-6 * np.linalg.norm(G["stgp_var"][i]), i


def test_gpytorch_expansion_and_difference_form_agree():
    # the centred quadratic expansion GPyTorch uses vs the difference form the CUDA kernel evaluates
    x, _ = orc.synth_field_data(300, seed=5)
    spec = orc.battgp_spec()
    a = orc.cov(spec, x, x)
    b = orc.cov(spec, x, x, gpytorch_expansion=True)
    assert np.max(np.abs(a - b)) < 1e-12 * np.max(np.abs(a))


@pytest.mark.parametrize("maker", [orc.battgp_spec, orc.matern_periodic_spec, lambda: orc.scaled_rbf_spec(4, 0.01, 20.0)])
def test_lml_gradient_matches_finite_differences(maker):
    # LML / gradient are unpinned by the reference: self-validate the oracle
    spec = maker()
    x, y = orc.synth_field_data(120, seed=9)
    noise = 2.33e-6
    g = orc.lml_grad(spec, x, y, noise)

    def lml(sp, nz):
        return orc.fit(sp, x, y, nz).lml

    h = 1e-6
    fd = (lml(spec, noise * (1 + h)) - lml(spec, noise * (1 - h))) / (2 * noise * h)
    assert abs(fd - g["noise"]) < 1e-4 * max(1.0, abs(g["noise"]))
    for ti, t in enumerate(spec.terms):
        import copy
        for attr, key in (("outputscale", "outputscale"),):
            sp, sm = copy.deepcopy(spec), copy.deepcopy(spec)
            v = getattr(t, attr)
            setattr(sp.terms[ti], attr, v * (1 + h)); setattr(sm.terms[ti], attr, v * (1 - h))
            fd = (lml(sp, noise) - lml(sm, noise)) / (2 * v * h)
            assert abs(fd - g["terms"][ti][key]) < 2e-4 * max(1.0, abs(fd)), (ti, key)
        for attr in ("lengthscale", "period"):
            vals = list(getattr(t, attr))
            for j in range(len(g["terms"][ti][attr])):
                sp, sm = copy.deepcopy(spec), copy.deepcopy(spec)
                vp, vm = list(vals), list(vals)
                vp[j] *= 1 + h; vm[j] *= 1 - h
                setattr(sp.terms[ti], attr, tuple(vp)); setattr(sm.terms[ti], attr, tuple(vm))
                fd = (lml(sp, noise) - lml(sm, noise)) / (2 * vals[j] * h)
                assert abs(fd - g["terms"][ti][attr][j]) < 2e-4 * max(1.0, abs(fd)), (ti, attr, j)


def test_fit_identities():
    x, y = orc.synth_field_data(400, seed=2)
    spec = orc.battgp_spec()
    f = orc.fit(spec, x, y, 2.33e-6)
    k = orc.train_cov(spec, x, 2.33e-6)
    assert np.linalg.norm(f.L @ f.L.T - k) / np.linalg.norm(k) < 1e-14
    assert np.linalg.norm(k @ f.alpha - y) / np.linalg.norm(y) < 1e-9
    sign, ld = np.linalg.slogdet(k)
    assert sign > 0 and abs(ld - f.logdet) < 1e-8 * abs(ld)
    assert abs(orc.mll_loss(f, 400) + f.lml / 400) < 1e-15


def test_psd_safe_cholesky_jitter_sequence():
    k = np.ones((6, 6))                         # rank one: not PD
    L, jit = orc.psd_safe_cholesky(k)
    assert jit in (1e-8, 1e-7, 1e-6)
    with pytest.raises(np.linalg.LinAlgError):
        orc.psd_safe_cholesky(-np.eye(4))


REAL = np.load(os.path.join(os.path.dirname(__file__), "golden", "real_field_data.npz")) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "golden", "real_field_data.npz")) else None
REAL_KEYS = ("b14_c1", "b14_cpack", "b3_c5")


@pytest.mark.parametrize("key", REAL_KEYS)
def test_real_field_data_fixture_is_reproducible(key):
    """tests/golden/real_field_data.npz holds training sets produced by the reference's own data layer
    (batt_data.py:180-256 generateTrainingData on tests/data/cache/*.feather) plus the oracle's results on them; the
    committed results must be what the oracle computes today, and the data must look like BattGP's (4 columns
    age/I/SOC/T, ages sorted, R > 0, every op-point feature inside the reference's filter ranges)."""
    assert REAL is not None
    x, y, xq = REAL[f"{key}_x"], REAL[f"{key}_y"], REAL[f"{key}_xq"]
    assert x.shape[1] == 4 and x.shape[0] == y.shape[0] and xq.shape == (300, 4)
    assert np.all(np.diff(x[:, 0]) >= 0) and np.all(np.isfinite(x)) and np.all(np.isfinite(y))
    th = REAL["theta"]
    spec = orc.battgp_spec(th[1], th[2], th[3:6])
    f = orc.fit(spec, x, y, th[0])
    mean, var = orc.predict(spec, x, f, xq)
    assert f.jitter == float(REAL[f"{key}_jitter"])
    assert abs(f.lml - float(REAL[f"{key}_lml"])) < 1e-9 * abs(f.lml)
    np.testing.assert_allclose(mean, REAL[f"{key}_mean"], rtol=1e-8)
    np.testing.assert_allclose(var, REAL[f"{key}_var"], rtol=1e-6)


def test_full_size_real_fixture_leading_block_is_reproducible():
    """real_field_data_40k.npz (all 40 000 filtered rows of system 14 / cell 1, oracle results from a one-off LAPACK run on
    the build host): the leading principal block of L depends only on the leading block of K, so the oracle at n0 = 1500
    must reproduce the stored diagonal of the full-size factor; logdet must equal 2 sum log diag."""
    R = np.load(os.path.join(os.path.dirname(__file__), "golden", "real_field_data_40k.npz"))
    x, y, th = R["x"], R["y"], R["theta"]
    assert x.shape == (40000, 4) and y.shape == (40000,) and np.all(np.diff(x[:, 0]) >= 0)
    assert abs(2.0 * np.log(R["l_diag"]).sum() - float(R["logdet"])) < 1e-6
    n0 = 1500
    spec = orc.battgp_spec(th[1], th[2], th[3:6])
    L = orc.cholesky_lower(orc.train_cov(spec, x[:n0], th[0]))
    np.testing.assert_allclose(np.diag(L), R["l_diag"][:n0], rtol=1e-9)
