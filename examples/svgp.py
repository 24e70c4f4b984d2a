"""The reference's examples/svgp.py on this package: a neural network feeding a sparse
variational GP classifier (SVGP, MultiClass likelihood with the robust-max link, one latent GP
per class, whiten=False), network and GP trained end to end by Adam on minibatches.

What changed relative to the reference script (examples/svgp.py:100-190): the network is a
torch module instead of a tf template, the placeholders / feed_dict / session are gone -- each
step assigns the minibatch's network features and labels to `gp_model.X` / `gp_model.Y` and
reads `gp_model.objective` -- and MNIST is replaced by synthetic 10-class data of the same layout
(flattened 28 x 28 images, integer labels in an [N, 1] column): there is no network access here.
The gradient w.r.t. the network reaches it through the Gram kernels' d/dX path (csrc/gram.cu).

    python examples/svgp.py [--iters 300]
"""
import argparse
import os.path as osp
import sys

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, osp.join(ROOT, 'gpflow-slim_b200'))

import gpflowSlim as gpf  # noqa: E402


def synthetic_mnist(n_train=6000, n_test=1000, num_classes=10, seed=100):
    """Class-dependent blobs rendered into 28 x 28 'images' (flattened, in [0, 1])."""
    rng = np.random.RandomState(seed)
    protos = rng.rand(num_classes, 784) < 0.15

    def draw(n):
        lab = rng.randint(0, num_classes, size=n)
        x = np.clip(protos[lab] * rng.uniform(0.6, 1.0, size=(n, 784)) + 0.25 * rng.rand(n, 784) *
                    (rng.rand(n, 784) < 0.2), 0.0, 1.0)
        return x.astype(np.float64), lab.astype(np.float64)[:, None]
    return draw(n_train) + draw(n_test)


class SmallNet(torch.nn.Module):
    """Stand-in for make_small_mnist_nn (examples/svgp.py:60-80): two conv + one dense layer,
    `end_h` features out, computed in float64 like the GP."""

    def __init__(self, end_h):
        super().__init__()
        self.c1 = torch.nn.Conv2d(1, 8, 5, stride=2, padding=2)
        self.c2 = torch.nn.Conv2d(8, 16, 5, stride=2, padding=2)
        self.fc = torch.nn.Linear(16 * 7 * 7, end_h)

    def forward(self, x_flat):
        x = x_flat.reshape(-1, 1, 28, 28)
        x = torch.relu(self.c1(x))
        x = torch.relu(self.c2(x))
        return self.fc(x.reshape(x.shape[0], -1))


def suggest_sensible_lengthscale(h):
    """Median-ish pairwise distance heuristic (examples/svgp.py:84-90)."""
    h = h.detach().cpu().numpy()
    sub = h[:500]
    d = np.sqrt(((sub[:, None, :] - sub[None, :, :]) ** 2).sum(-1))
    return float(np.mean(d)) * np.ones(h.shape[1])


def suggest_good_initial_inducing_points(h, num_inducing, seed=0):
    """The reference runs k-means on network features (examples/svgp.py:92-97); a random subset
    of those features serves the same purpose."""
    h = h.detach().cpu().numpy()
    idx = np.random.RandomState(seed).permutation(h.shape[0])[:num_inducing]
    return h[idx].copy()


def main(iters=300, report=50, num_h=32, num_inducing=100, minibatch_size=250, quiet=False,
         n_train=6000, n_test=1000):
    dev = gpf.settings.device
    x_train, y_train, x_test, y_test = synthetic_mnist(n_train, n_test)
    to = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    x_train_t, y_train_t, x_test_t, y_test_t = map(to, (x_train, y_train, x_test, y_test))
    num_classes = 10

    torch.manual_seed(0)
    nn_base = SmallNet(num_h).to(device=dev, dtype=torch.float64)
    with torch.no_grad():
        h0 = nn_base(x_train_t[:2000])

    kernel = gpf.kernels.RBF(num_h, lengthscales=suggest_sensible_lengthscale(h0), ARD=True)
    likelihood = gpf.likelihoods.MultiClass(num_classes)
    gp_model = gpf.models.SVGP(h0[:minibatch_size], y_train_t[:minibatch_size], kernel, likelihood,
                               Z=suggest_good_initial_inducing_points(h0, num_inducing),
                               num_latent=num_classes, whiten=False, minibatch_size=None,
                               num_data=x_train.shape[0])

    def test_metrics():
        with torch.no_grad():
            h = nn_base(x_test_t)
            fmu, fvar = gp_model._build_predict(h)
            ll = gp_model.likelihood.predict_density(fmu, fvar, y_test_t).mean()
            prob, _ = gp_model.likelihood.predict_mean_and_var(fmu, fvar)
            acc = (prob.argmax(1) == y_test_t.squeeze().long()).double().mean()
        return float(acc), float(ll)

    optimiser = gpf.training.AdamOptimizer()          # tf.train.AdamOptimizer() defaults
    var_list = gp_model.trainable_tensors + list(nn_base.parameters())
    data_indx, loss = 0, None
    for i in range(iters):
        indx = np.mod(np.arange(data_indx, data_indx + minibatch_size), x_train.shape[0])
        data_indx += minibatch_size
        gp_model.X = nn_base(x_train_t[indx])
        gp_model.Y = y_train_t[indx]
        loss = -float(optimiser.minimize(gp_model, var_list=var_list))
        if i % report == 0 and not quiet:
            acc, ll = test_metrics()
            print('Iteration {}: Loss is {}. \nTest set LL {}, Acc {}'.format(i, loss, ll, acc))
    acc, ll = test_metrics()
    if not quiet:
        print('final: Loss {}  test LL {}  Acc {}'.format(loss, ll, acc))
    return loss, acc, ll


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=300)
    ap.add_argument('--num_h', type=int, default=32,
                    help='network features fed to the GP (the reference uses 100; above 32 the Gram '
                         'matrices take the composed GEMM path instead of the fused kernel)')
    a = ap.parse_args()
    main(iters=a.iters, num_h=a.num_h)
