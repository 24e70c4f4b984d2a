"""Known-answer tests of oracle/tf_shim against the DOCUMENTED behaviour of the TensorFlow 1.x
ops the reference's hot path calls -- the worked examples of the TF API documentation
(tf.tile, tf.gather, tf.scatter_nd, tf.one_hot, tf.matrix_band_part, tf.matrix_diag,
tf.reduce_sum ...) and the documented contracts of tf.cholesky / tf.matrix_triangular_solve /
tf.clip_by_value.  The golden vectors under tests/golden/ come from the UNMODIFIED reference
running over this shim; these tests are the evidence that the shim means what TensorFlow means."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope='module')
def tf():
    # imported under a private name: `tensorflow` must not leak into other tests' sys.modules
    import importlib.util
    path = os.path.join(ROOT, 'oracle', 'tf_shim', 'tensorflow', '__init__.py')
    spec = importlib.util.spec_from_file_location('_tf_shim_under_test', path)
    mod = importlib.util.module_from_spec(spec)
    old_default = torch.get_default_dtype()
    spec.loader.exec_module(mod)
    yield mod
    torch.set_default_dtype(old_default)


def eq(a, b):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=0, atol=1e-14)


def test_shape_ops_follow_the_tf_docs(tf):
    a = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float64)
    eq(tf.tile(a, [1, 2]), [[1, 2, 3, 1, 2, 3], [4, 5, 6, 4, 5, 6]])           # tf.tile doc example
    eq(tf.tile(a, [2, 1]), [[1, 2, 3], [4, 5, 6], [1, 2, 3], [4, 5, 6]])
    eq(tf.transpose(a), a.T)
    x3 = np.arange(24.0).reshape(2, 3, 4)
    eq(tf.transpose(x3, [2, 0, 1]), x3.transpose(2, 0, 1))
    eq(tf.matrix_transpose(x3), x3.transpose(0, 2, 1))
    assert tuple(tf.expand_dims(a, 1).shape) == (2, 1, 3)
    assert tuple(tf.squeeze(np.zeros((1, 2, 1, 3))).shape) == (2, 3)
    assert tuple(tf.squeeze(np.zeros((1, 2, 1, 3)), 2).shape) == (1, 2, 3)
    eq(tf.concat([a, a], 0), np.concatenate([a, a], 0))
    eq(tf.stack([a[0], a[1]], axis=1), np.stack([a[0], a[1]], 1))
    assert tf.shape(a).tolist() == [2, 3] and int(tf.size(a)) == 6 and tf.rank(a) == 2
    eq(tf.reshape(a, [3, -1]), a.reshape(3, -1))
    eq(tf.fill([2, 3], 9.0), np.full((2, 3), 9.0))
    eq(tf.eye(3, dtype=np.float64), np.eye(3))


def test_gather_scatter_one_hot_follow_the_tf_docs(tf):
    p = np.array([10., 11., 12., 13., 14., 15.])
    eq(tf.gather(p, [2, 0, 2, 5]), [12, 10, 12, 15])
    eq(tf.gather(p, np.array([[1, 2], [0, 3]])), [[11, 12], [10, 13]])        # indices.shape is kept
    m = np.arange(12.0).reshape(4, 3)
    eq(tf.gather(m, [3, 1]), m[[3, 1]])
    eq(tf.gather(m, [2, 0], axis=1), m[:, [2, 0]])
    # tf.scatter_nd doc example
    eq(tf.scatter_nd(np.array([[4], [3], [1], [7]]), np.array([9., 10., 11., 12.]), [8]),
       [0, 11, 0, 10, 9, 0, 0, 12])
    # tf.one_hot doc examples
    eq(tf.one_hot(np.array([0, 1, 2]), 3), np.eye(3))
    eq(tf.one_hot(np.array([0, 2, 1]), 3, on_value=5.0, off_value=0.0), [[5, 0, 0], [0, 0, 5], [0, 5, 0]])
    assert tf.argmax(np.array([[1., 9., 3.], [7., 2., 8.]]), axis=1).tolist() == [1, 2]
    assert tf.argmax(np.array([[1., 9., 3.], [7., 2., 8.]]), axis=0).tolist() == [1, 0, 1]


def test_reductions_follow_the_tf_docs(tf):
    x = np.array([[1., 1., 1.], [1., 1., 1.]])
    eq(tf.reduce_sum(x), 6)                                                     # tf.reduce_sum doc example
    eq(tf.reduce_sum(x, 0), [2, 2, 2])
    eq(tf.reduce_sum(x, 1), [3, 3])
    eq(tf.reduce_sum(x, 1, keepdims=True), [[3], [3]])
    eq(tf.reduce_sum(x, 1, keep_dims=True), [[3], [3]])                         # TF-1.x spelling
    eq(tf.reduce_sum(x, [0, 1]), 6)
    y = np.array([[1., 2.], [3., 4.]])
    eq(tf.reduce_mean(y), 2.5)
    eq(tf.reduce_mean(y, 0), [2, 3])
    eq(tf.reduce_prod(y, 1), [2, 12])
    eq(tf.reduce_max(y, 0), [3, 4])
    eq(tf.trace(y), 5)
    eq(tf.add_n([y, y, y]), 3 * y)
    eq(tf.norm(y), np.sqrt(30.0))


def test_band_part_and_diag_follow_the_tf_docs(tf):
    inp = np.array([[0, 1, 2, 3], [-1, 0, 1, 2], [-2, -1, 0, 1], [-3, -2, -1, 0]], dtype=np.float64)
    eq(tf.matrix_band_part(inp, -1, 0), np.tril(inp))                           # "Lower triangular part"
    eq(tf.matrix_band_part(inp, 0, -1), np.triu(inp))                           # "Upper triangular part"
    d = np.array([[1., 2., 3., 4.], [5., 6., 7., 8.]])
    out = tf.matrix_diag(d)                                                     # tf.matrix_diag doc example
    assert tuple(out.shape) == (2, 4, 4)
    eq(out[0], np.diag(d[0]))
    eq(out[1], np.diag(d[1]))
    eq(tf.matrix_diag_part(out), d)
    eq(tf.diag_part(np.diag([1., 2., 3.])), [1, 2, 3])


def test_linear_algebra_contracts(tf):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((5, 5))
    S = A @ A.T + 5 * np.eye(5)
    L = np.linalg.cholesky(S)
    # tf.cholesky: lower factor; "only the lower-triangular part of the input will be used"
    junk = np.tril(S) + np.triu(rng.standard_normal((5, 5)), 1)
    eq(tf.cholesky(S), L)
    np.testing.assert_allclose(tf.cholesky(junk).numpy(), L, atol=1e-13)
    B = rng.standard_normal((5, 3))
    # tf.matrix_triangular_solve: lower=True by default, the other triangle "is assumed to be zero
    # and not accessed"; adjoint=True solves with the adjoint of that triangle
    Lj = L + np.triu(rng.standard_normal((5, 5)), 1)
    np.testing.assert_allclose(tf.matrix_triangular_solve(Lj, B).numpy(), np.linalg.solve(L, B), atol=1e-13)
    np.testing.assert_allclose(tf.matrix_triangular_solve(Lj, B, lower=True, adjoint=True).numpy(),
                               np.linalg.solve(L.T, B), atol=1e-13)
    U = L.T + np.tril(rng.standard_normal((5, 5)), -1)
    np.testing.assert_allclose(tf.matrix_triangular_solve(U, B, lower=False).numpy(),
                               np.linalg.solve(L.T, B), atol=1e-13)
    np.testing.assert_allclose(tf.matrix_inverse(S).numpy(), np.linalg.inv(S), atol=1e-13)
    C = rng.standard_normal((4, 5))
    eq(tf.matmul(C, A), C @ A)
    eq(tf.matmul(C, C, transpose_b=True), C @ C.T)
    eq(tf.matmul(C, C, transpose_a=True), C.T @ C)
    batch = rng.standard_normal((3, 4, 5))
    eq(tf.matmul(batch, batch, transpose_b=True), batch @ batch.transpose(0, 2, 1))
    eq(tf.einsum('ij,jk->ik', C, A), C @ A)


def test_elementwise_and_gradient_contracts(tf):
    x = np.array([-30.0, -1.0, 0.0, 2.0, 40.0])
    np.testing.assert_allclose(tf.nn.softplus(x).numpy(), np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0),
                               rtol=1e-14)                                      # log(exp(x) + 1), no cutoff
    eq(tf.nn.relu(x), np.maximum(x, 0))
    eq(tf.square(x), x ** 2)
    eq(tf.where(torch.tensor([True, False, True]), np.array([1., 2., 3.]), np.array([9., 8., 7.])), [1, 8, 3])
    eq(tf.equal(np.array([1., 2.]), np.array([1., 3.])).to(torch.float64), [1, 0])
    eq(tf.cast(np.array([1.7, -1.7]), np.int32), [1, -1])                       # truncation towards zero
    eq(tf.pow(np.array([2., 3.]), np.array([3., 2.])), [8, 9])
    # tf.clip_by_value: gradient passes where the value is not clipped (incl. the boundary), 0 elsewhere
    v = torch.tensor([-1.0, 0.0, 0.5, 2.0], requires_grad=True)
    out = tf.clip_by_value(v, 0.0, np.inf)
    eq(out, [0, 0, 0.5, 2])
    (g,) = torch.autograd.grad(out.sum(), [v])
    eq(g, [0, 1, 1, 1])
    # variables: tf.get_variable creates trainable float64 leaves in creation order
    tf.shim_reset()
    a = tf.get_variable('a', initializer=np.array([1.0, 2.0]))
    b = tf.get_variable('b', initializer=np.array(3.0), trainable=False)
    assert [t.tf_name for t in tf.shim_variables()] == ['a', 'b']
    assert a.requires_grad and not b.requires_grad and a.dtype == torch.float64
    a.assign(np.array([5.0, 6.0]))
    eq(a, [5, 6])
    tf.shim_reset()
