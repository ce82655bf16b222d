"""sklearn kernel object -> device covariance spec.

The reference hands ``Kriging`` an arbitrary ``sklearn.gaussian_process.kernels``
expression (kriging.py:80-96, driver.py:181-190).  The device kernels evaluate

    k(x, x') = c0 + c1 * stat(||(x - x') / ls||),     kdiag = kernel_.diag(x)

which covers every tree the reference's tests and docs use (SURVEY.md §8 A5
note): ``ConstantKernel * Matern`` (default, kriging.py:85-88), ``Matern()``
alone, ``ConstantKernel + Matern``, each optionally ``+ WhiteKernel``; with
Matern nu in {0.5, 1.5, 2.5, inf} or RBF, isotropic or ARD.  Anything else
raises ``NotImplementedError`` -- there is no silent fallback.

sklearn semantics restated: ``Sum`` adds, ``Product`` multiplies
(sklearn:kernels.py:838-871, 936-990); ``WhiteKernel`` is 0 off the training
diagonal and ``noise_level`` on ``diag`` (kernels.py:1417-1425); stationary
kernels have unit ``diag``.
"""
import numpy as np
from sklearn.gaussian_process import kernels as sk

KIND_MATERN12, KIND_MATERN32, KIND_MATERN52, KIND_RBF, KIND_MATERN_INF = range(5)


class KernelSpec:
    __slots__ = ('kind', 'c0', 'c1', 'kdiag', 'ls')

    def __init__(self, kind, c0, c1, kdiag, ls):
        self.kind, self.c0, self.c1, self.kdiag, self.ls = kind, c0, c1, kdiag, ls

    def __repr__(self):
        return "KernelSpec(kind=%d, c0=%r, c1=%r, kdiag=%r, ls=%r)" % (
            self.kind, self.c0, self.c1, self.kdiag, self.ls)


def _stationary_factor(kernel, dim):
    ls = np.atleast_1d(np.asarray(kernel.length_scale, dtype=np.float64))
    if ls.size == 1:
        ls = np.full(dim, float(ls[0]))
    elif ls.size != dim:
        raise ValueError("Anisotropic kernel must have the same number of dimensions as data (%d!=%d)"
                         % (ls.size, dim))   # same message as sklearn's _check_length_scale
    if isinstance(kernel, sk.Matern):
        nu = float(kernel.nu)
        kind = {0.5: KIND_MATERN12, 1.5: KIND_MATERN32, 2.5: KIND_MATERN52, np.inf: KIND_MATERN_INF}.get(nu)
        if kind is None:
            raise NotImplementedError("Matern(nu=%r): only nu in {0.5, 1.5, 2.5, inf} has a device kernel" % nu)
    else:
        kind = KIND_RBF
    return (kind, tuple(ls.tolist()))


def _cross_poly(kernel, dim):
    """k(x, x') for x != x' as {tuple(sorted factors): coefficient}."""
    if isinstance(kernel, sk.ConstantKernel):
        return {(): float(kernel.constant_value)}
    if isinstance(kernel, sk.WhiteKernel):
        return {}
    if isinstance(kernel, sk.RBF):       # Matern subclasses RBF
        return {(_stationary_factor(kernel, dim),): 1.0}
    if isinstance(kernel, sk.Sum):
        out = dict(_cross_poly(kernel.k1, dim))
        for key, coef in _cross_poly(kernel.k2, dim).items():
            out[key] = out.get(key, 0.0) + coef
        return out
    if isinstance(kernel, sk.Product):
        out = {}
        for ka, ca in _cross_poly(kernel.k1, dim).items():
            for kb, cb in _cross_poly(kernel.k2, dim).items():
                key = tuple(sorted(ka + kb))
                out[key] = out.get(key, 0.0) + ca * cb
        return out
    raise NotImplementedError("kernel %r has no device implementation (supported: Sum/Product of "
                              "ConstantKernel, WhiteKernel, RBF, Matern)" % (kernel,))


def _diag(kernel):
    if isinstance(kernel, sk.ConstantKernel):
        return float(kernel.constant_value)
    if isinstance(kernel, sk.WhiteKernel):
        return float(kernel.noise_level)
    if isinstance(kernel, sk.RBF):
        return 1.0
    if isinstance(kernel, sk.Sum):
        return _diag(kernel.k1) + _diag(kernel.k2)
    if isinstance(kernel, sk.Product):
        return _diag(kernel.k1) * _diag(kernel.k2)
    raise NotImplementedError("kernel %r has no device implementation" % (kernel,))


def flatten(kernel, dim):
    """-> KernelSpec, or NotImplementedError if the tree is not affine in one stationary factor."""
    poly = {k: c for k, c in _cross_poly(kernel, dim).items() if c != 0.0}
    factors = sorted({f for key in poly for f in key})
    if len(factors) > 1 or any(len(key) > 1 for key in poly):
        raise NotImplementedError(
            "kernel %r mixes several stationary factors; the device path evaluates "
            "c0 + c1 * stat(r) with ONE RBF/Matern factor" % (kernel,))
    c0 = poly.get((), 0.0)
    if factors:
        kind, ls = factors[0]
        c1 = poly.get((factors[0],), 0.0)
    else:  # pure constant (+ white): any kind with c1 = 0
        kind, ls, c1 = KIND_RBF, (1.0,) * dim, 0.0
    return KernelSpec(kind, float(c0), float(c1), _diag(kernel), np.asarray(ls, dtype=np.float64))
