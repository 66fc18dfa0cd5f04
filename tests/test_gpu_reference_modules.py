"""GPU: the reference's OWN modules, unmodified, on the import shim: BatteryCellGP_Full
.predict / .predict_r0_op / .train_hyperparameters, the three trainers of src/gp/training.py, ScaledRBFModel and the
reference's tests/gp unit tests.  The checks live in tools/run_reference_modules.py; they need a reference tree, which is
/root/reference in the build container and the staged oracle/_ref/reference_src.tar.gz (tools/stage_reference.py,
git-ignored) on the GPU box -- skipped when neither is there."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_reference_modules as rrm  # noqa: E402


@pytest.fixture(scope="module")
def ref(eng):
    import torch
    path = rrm.find_reference()
    if path is None:
        pytest.skip("no reference tree staged (tools/stage_reference.py)")
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)          # gp_runner.py:158-159
    rrm.install(path)
    yield path
    torch.set_default_dtype(prev)
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    if path in sys.path:
        sys.path.remove(path)


def _run(fn, *a):
    import torch
    res = []
    ok = fn(*a, res.append) if a else fn(torch.device("cuda", 0), res.append)
    bad = [r for r in res if r.get("ok") is False]
    assert ok and not bad, bad


def test_reference_battcellgp_full_predicts_the_golden_results(ref):
    _run(rrm.check_predict)


def test_reference_trainers_follow_the_oracle_trajectory(ref):
    _run(rrm.check_training)


def test_reference_scaled_rbf_model(ref):
    _run(rrm.check_scaled_rbf)


def test_reference_own_unit_tests_pass_on_the_shim(ref):
    res = []
    ok = rrm.run_reference_unit_tests(ref, res.append)
    assert ok, [r for r in res if r.get("ok") is False]
