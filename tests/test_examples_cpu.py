"""examples/gpr.py and examples/svgp.py (the reference's two example scripts on this package)
run a few steps on the torch-CPU test double of the ops layer: the scripts' host logic --
minibatch reassignment of X / Y, joint Adam over network and GP tensors, evaluation -- works and
the objective goes down.  The same scripts run on the real kernels in test_gpu_zz_widened.py."""
import importlib.util
import os

import pytest

import cpu_ops_double

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location('example_' + name, os.path.join(ROOT, 'examples', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture
def on_cpu_double(monkeypatch):
    import gpflowSlim
    cpu_ops_double.install(monkeypatch)
    old = gpflowSlim.settings.device
    gpflowSlim.settings.device = 'cpu'
    yield
    gpflowSlim.settings.device = None if old.type == 'cpu' else old


def test_gpr_example_adam_and_lbfgs(on_cpu_double):
    ex = _load('gpr')
    first, _, _ = ex.main(iters=1, quiet=True)
    last, rmse, ll = ex.main(iters=25, quiet=True)
    assert last < first
    obj, rmse2, ll2 = ex.main(lbfgs=True, quiet=True)
    assert obj < last and rmse2 < 6.0 and ll2 > -3.5


def test_svgp_example_trains(on_cpu_double):
    ex = _load('svgp')
    loss, acc, ll = ex.main(iters=12, quiet=True, n_train=600, n_test=200, num_inducing=20,
                            minibatch_size=100, num_h=8)
    assert loss == loss and 0.0 <= acc <= 1.0 and ll == ll      # finite
