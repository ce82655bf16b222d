"""numpy restatement of the scikit-learn GP arithmetic behind ``Kriging.evaluate``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (`/root/reference/batman/surrogate/kriging.py:117-123, 209-212`)
delegates all arithmetic to ``sklearn.gaussian_process.GaussianProcessRegressor``
(pinned ``scikit-learn==0.23.1`` in requirements.txt:51; 1.9.0 is installed
here and on the GPU box).  ``sklearn:`` citations below refer to
``site-packages/sklearn/gaussian_process/`` of that 1.9.0 install.

A kernel is a nested tuple tree::

    ('const', c) | ('white', noise) | ('matern', nu, ls) | ('rbf', ls)
    ('sum', a, b) | ('prod', a, b)

which is the subset the reference exercises (SURVEY.md §8 A5 note).
"""
import math

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular

JITTER = 1e-10  # GaussianProcessRegressor(alpha=1e-10) default, sklearn:_gpr.py:215


# --------------------------------------------------------------------------
# kernel tree
# --------------------------------------------------------------------------
def from_sklearn(kernel):
    """Convert an sklearn kernel object into the tuple tree."""
    from sklearn.gaussian_process import kernels as sk

    if isinstance(kernel, sk.ConstantKernel):
        return ('const', float(kernel.constant_value))
    if isinstance(kernel, sk.WhiteKernel):
        return ('white', float(kernel.noise_level))
    if isinstance(kernel, sk.Matern):  # Matern subclasses RBF: test it first
        return ('matern', float(kernel.nu),
                np.atleast_1d(np.asarray(kernel.length_scale, dtype=np.float64)))
    if isinstance(kernel, sk.RBF):
        return ('rbf', np.atleast_1d(np.asarray(kernel.length_scale, dtype=np.float64)))
    if isinstance(kernel, sk.Sum):
        return ('sum', from_sklearn(kernel.k1), from_sklearn(kernel.k2))
    if isinstance(kernel, sk.Product):
        return ('prod', from_sklearn(kernel.k1), from_sklearn(kernel.k2))
    raise NotImplementedError("kernel %r is outside the hot-path subset" % (kernel,))


def _scaled_dists(Xa, Xb, ls, dtype=np.float64):
    """cdist(Xa / ls, Xb / ls, 'euclidean') -- sklearn:kernels.py:1716-1721.

    scipy's C loop is a sequential, non-fused ``s += d*d`` over the columns
    followed by one sqrt (probed bit-for-bit in tests/test_oracle_gp.py).
    """
    Xa = np.asarray(Xa, dtype=dtype)
    Xb = np.asarray(Xb, dtype=dtype)
    ls = np.asarray(ls, dtype=dtype)
    U = Xa / ls
    V = Xb / ls
    s = np.zeros((U.shape[0], V.shape[0]), dtype=dtype)
    for c in range(U.shape[1]):
        d = U[:, c, None] - V[None, :, c]
        s = s + d * d
    return np.sqrt(s)


def _stationary(kind, nu, dists, dtype=np.float64):
    """sklearn:kernels.py:1723-1730 (Matern) and 1557-1563 (RBF)."""
    one = dtype(1.0)
    if kind == 'rbf' or nu == np.inf:
        # RBF: exp(-0.5 * sqeuclidean); Matern(nu=inf): exp(-dists**2 / 2)
        return np.exp(-(dists ** 2) / dtype(2.0))
    if nu == 0.5:
        return np.exp(-dists)
    if nu == 1.5:
        K = dists * dtype(math.sqrt(3)) if dtype is np.float64 else dists * np.sqrt(dtype(3))
        return (one + K) * np.exp(-K)
    if nu == 2.5:
        K = dists * dtype(math.sqrt(5)) if dtype is np.float64 else dists * np.sqrt(dtype(5))
        return (one + K + K ** 2 / dtype(3.0)) * np.exp(-K)
    raise NotImplementedError("generic-nu Matern (Bessel K) is outside the hot-path subset")


def kernel_cross(tree, Xa, Xb, dtype=np.float64):
    """k(Xa, Xb) for distinct sets (WhiteKernel contributes 0, kernels.py:1417-1418)."""
    tag = tree[0]
    shape = (len(Xa), len(Xb))
    if tag == 'const':
        return np.full(shape, tree[1], dtype=dtype)  # kernels.py:1290-1295
    if tag == 'white':
        return np.zeros(shape, dtype=dtype)
    if tag == 'matern':
        return _stationary('matern', tree[1], _scaled_dists(Xa, Xb, tree[2], dtype), dtype)
    if tag == 'rbf':
        if dtype is np.float64:
            # RBF uses sqeuclidean directly: cdist(..., 'sqeuclidean'), kernels.py:1557-1563
            U = np.asarray(Xa, dtype) / tree[1]
            V = np.asarray(Xb, dtype) / tree[1]
            s = np.zeros(shape)
            for c in range(U.shape[1]):
                d = U[:, c, None] - V[None, :, c]
                s = s + d * d
            return np.exp(-0.5 * s)
        d = _scaled_dists(Xa, Xb, tree[1], dtype)
        return np.exp(-(d * d) / dtype(2.0))
    if tag == 'sum':
        return kernel_cross(tree[1], Xa, Xb, dtype) + kernel_cross(tree[2], Xa, Xb, dtype)  # kernels.py:870
    if tag == 'prod':
        return kernel_cross(tree[1], Xa, Xb, dtype) * kernel_cross(tree[2], Xa, Xb, dtype)  # kernels.py:968
    raise ValueError(tag)


def kernel_train(tree, X, dtype=np.float64):
    """k(X, X): stationary kernels have an exact unit diagonal (squareform +
    fill_diagonal, kernels.py:1741-1744), WhiteKernel is noise * I."""
    tag = tree[0]
    n = len(X)
    if tag == 'white':
        return dtype(tree[1]) * np.eye(n, dtype=dtype)
    if tag in ('matern', 'rbf'):
        K = kernel_cross(tree, X, X, dtype)
        np.fill_diagonal(K, dtype(1.0))
        return K
    if tag == 'sum':
        return kernel_train(tree[1], X, dtype) + kernel_train(tree[2], X, dtype)
    if tag == 'prod':
        return kernel_train(tree[1], X, dtype) * kernel_train(tree[2], X, dtype)
    return kernel_cross(tree, X, X, dtype)


def kernel_diag(tree):
    """kernel_.diag(X): a constant for this kernel family (kernels.py:873-890, 971-990)."""
    tag = tree[0]
    if tag in ('const', 'white'):
        return float(tree[1])
    if tag in ('matern', 'rbf'):
        return 1.0
    if tag == 'sum':
        return kernel_diag(tree[1]) + kernel_diag(tree[2])
    if tag == 'prod':
        return kernel_diag(tree[1]) * kernel_diag(tree[2])
    raise ValueError(tag)


# --------------------------------------------------------------------------
# fit tail / predict / LML
# --------------------------------------------------------------------------
def normalise_y(y):
    """normalize_y=True branch, sklearn:_gpr.py:275-280."""
    y = np.asarray(y, dtype=np.float64)
    mean = np.mean(y, axis=0)
    std = np.std(y, axis=0)
    if std == 0.0 or std < 10 * np.finfo(np.float64).eps:  # _handle_zeros_in_scale
        std = 1.0
    return (y - mean) / std, float(mean), float(std)


def fit_state(tree, X, y):
    """Final factorisation at fixed hyper-parameters, sklearn:_gpr.py:349-367."""
    X = np.asarray(X, dtype=np.float64)
    yn, y_mean, y_std = normalise_y(y)
    K = kernel_train(tree, X)
    K[np.diag_indices_from(K)] += JITTER
    L = cholesky(K, lower=True, check_finite=False)
    alpha = cho_solve((L, True), yn, check_finite=False)
    return dict(tree=tree, X_train=X, alpha=alpha, L=L, y_mean=y_mean, y_std=y_std)


def state_from_sklearn(gp):
    """Fitted GaussianProcessRegressor -> oracle state (no arithmetic)."""
    return dict(tree=from_sklearn(gp.kernel_), X_train=np.asarray(gp.X_train_, dtype=np.float64),
                alpha=np.asarray(gp.alpha_, dtype=np.float64).ravel(),
                L=np.asarray(gp.L_, dtype=np.float64),
                y_mean=float(np.ravel(gp._y_train_mean)[0]) if np.ndim(gp._y_train_mean) else float(gp._y_train_mean),
                y_std=float(np.ravel(gp._y_train_std)[0]) if np.ndim(gp._y_train_std) else float(gp._y_train_std))


def predict(state, Xs, return_std=True):
    """GaussianProcessRegressor.predict(return_std=True), sklearn:_gpr.py:444-500."""
    Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
    K_trans = kernel_cross(state['tree'], Xs, state['X_train'])           # :446
    y_mean = K_trans @ state['alpha']                                      # :447
    y_mean = state['y_std'] * y_mean + state['y_mean']                     # :450
    if not return_std:
        return y_mean
    V = solve_triangular(state['L'], K_trans.T, lower=True, check_finite=False)   # :460
    y_var = np.full(len(Xs), kernel_diag(state['tree']))                   # :480
    y_var -= np.einsum("ij,ji->i", V.T, V)                                 # :481
    y_var[y_var < 0] = 0.0                                                 # :485-491
    y_var = y_var * state['y_std'] ** 2                                    # :494
    return y_mean, np.sqrt(y_var)                                          # :500


def evaluate(states, points):
    """Kriging.evaluate over a list of per-output states: (k, m) mean and sigma.

    Follows kriging.py:192-214 + functions/utils.py:19-31 (shape contract), but
    batched over points: sklearn's own per-point and batched paths agree to
    ~1e-12, which is the reference's noise floor.
    """
    points = np.atleast_2d(points)
    mean = np.empty((len(points), len(states)))
    sigma = np.empty_like(mean)
    for q, st in enumerate(states):
        mean[:, q], sigma[:, q] = predict(st, points)
    return mean, sigma


def lml(tree, X, y_norm):
    """log_marginal_likelihood(theta) without gradient, sklearn:_gpr.py:584-617."""
    X = np.asarray(X, dtype=np.float64)
    K = kernel_train(tree, X)
    K[np.diag_indices_from(K)] += JITTER                                   # :589
    try:
        L = cholesky(K, lower=True, check_finite=False)                    # :591
    except np.linalg.LinAlgError:
        return -np.inf                                                     # :593
    alpha = cho_solve((L, True), y_norm, check_finite=False)               # :601
    val = -0.5 * float(y_norm @ alpha)                                     # :613
    val -= np.log(np.diag(L)).sum()                                        # :614
    val -= K.shape[0] / 2 * np.log(2 * np.pi)                              # :615
    return val
