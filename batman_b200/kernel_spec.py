"""sklearn kernel object -> device covariance spec.

The reference hands ``Kriging`` an arbitrary ``sklearn.gaussian_process.kernels``
expression (kriging.py:80-96, driver.py:181-190).  A Sum/Product tree over
ConstantKernel / WhiteKernel / RBF / Matern(nu in {0.5, 1.5, 2.5, inf}) leaves,
isotropic or ARD, is expanded to

    k(x, x') = sum_t coef_t * prod_f stat_f(||(x - x') / ls_f||)^e_tf,    kdiag = kernel_.diag(x)

* one stationary leaf, at most linear  ->  ``c0 + c1 * stat(r)``: the fast kernels
  (every tree the reference's tests and docs use, SURVEY.md §8 A5 note:
  ``ConstantKernel * Matern`` (default, kriging.py:85-88), ``Matern()`` alone,
  ``ConstantKernel + Matern``, each optionally ``+ WhiteKernel``);
* up to 3 distinct stationary leaves and 6 terms  ->  ``general`` spec, evaluated by
  the general kernels (csrc/gp_predict_gen.cu, k_build_K);
* anything else (other kernel classes, generic-nu Matern, larger expansions) raises
  ``NotImplementedError`` -- there is no silent fallback.

sklearn semantics restated: ``Sum`` adds, ``Product`` multiplies
(sklearn:kernels.py:838-871, 936-990); ``WhiteKernel`` is 0 off the training
diagonal and ``noise_level`` on ``diag`` (kernels.py:1417-1425); stationary
kernels have unit ``diag``.
"""
import numpy as np
from sklearn.gaussian_process import kernels as sk

KIND_MATERN12, KIND_MATERN32, KIND_MATERN52, KIND_RBF, KIND_MATERN_INF, KIND_GENERAL = range(6)
GEN_MAX_F, GEN_MAX_T, GEN_MAX_FP, GEN_MAX_EXP = 3, 6, 48, 4     # csrc/model.h


class GeneralSpec:
    """Expanded tree: leaves (fkind[f], fls[f]) and terms (tcoef[t], texp[t][f])."""
    __slots__ = ('fkind', 'fls', 'tcoef', 'texp')

    def __init__(self, fkind, fls, tcoef, texp):
        self.fkind = [int(k) for k in fkind]
        self.fls = np.asarray(fls, dtype=np.float64)            # (nf, p)
        self.tcoef = np.asarray(tcoef, dtype=np.float64)        # (nt,)
        self.texp = np.asarray(texp, dtype=np.int32)            # (nt, nf)

    @property
    def nf(self):
        return len(self.fkind)

    @property
    def nt(self):
        return len(self.tcoef)

    def shape_key(self):
        return (tuple(self.fkind), tuple(map(tuple, self.texp.tolist())))


class KernelSpec:
    __slots__ = ('kind', 'c0', 'c1', 'kdiag', 'ls', 'general')

    def __init__(self, kind, c0, c1, kdiag, ls, general=None):
        self.kind, self.c0, self.c1, self.kdiag, self.ls, self.general = kind, c0, c1, kdiag, ls, general

    @property
    def nonneg(self):
        """Every coefficient >= 0: k(x, x') >= 0, K* can be sliced into unsigned bytes (lml.int8_scheme)."""
        if self.general is not None:
            return bool(np.all(self.general.tcoef >= 0.0))
        return self.c0 >= 0.0 and self.c1 >= 0.0

    def __repr__(self):
        if self.general is not None:
            g = self.general
            return "KernelSpec(general, fkind=%r, fls=%r, tcoef=%r, texp=%r, kdiag=%r)" % (
                g.fkind, g.fls.tolist(), g.tcoef.tolist(), g.texp.tolist(), self.kdiag)
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


def _cross_terms(kernel, dim, leaves):
    """k(x, x') for x != x' as an ordered list of [coef, leaf indices in multiplication order]; ``leaves``
    collects the distinct stationary leaves in traversal order."""
    if isinstance(kernel, sk.ConstantKernel):
        return [[float(kernel.constant_value), ()]]
    if isinstance(kernel, sk.WhiteKernel):
        return []
    if isinstance(kernel, sk.RBF):       # Matern subclasses RBF
        leaf = _stationary_factor(kernel, dim)
        if leaf not in leaves:
            leaves.append(leaf)
        return [[1.0, (leaves.index(leaf),)]]
    if isinstance(kernel, sk.Sum):
        out = _cross_terms(kernel.k1, dim, leaves)
        for coef, idx in _cross_terms(kernel.k2, dim, leaves):
            for term in out:
                if sorted(term[1]) == sorted(idx):
                    term[0] += coef
                    break
            else:
                out.append([coef, idx])
        return out
    if isinstance(kernel, sk.Product):
        ta = _cross_terms(kernel.k1, dim, leaves)
        tb = _cross_terms(kernel.k2, dim, leaves)
        out = []
        for ca, ia in ta:
            for cb, ib in tb:
                idx = ia + ib
                for term in out:
                    if sorted(term[1]) == sorted(idx):
                        term[0] += ca * cb
                        break
                else:
                    out.append([ca * cb, idx])
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
    """-> KernelSpec (affine in one stationary leaf, or ``general``); NotImplementedError outside the device subset."""
    leaves = []
    terms = [t for t in _cross_terms(kernel, dim, leaves) if t[0] != 0.0]
    used = sorted({i for _, idx in terms for i in idx})
    kdiag = _diag(kernel)
    if len(used) <= 1 and all(len(idx) <= 1 for _, idx in terms):
        c0 = sum(c for c, idx in terms if not idx)
        c1 = sum(c for c, idx in terms if idx)
        if used:
            kind, ls = leaves[used[0]]
        else:  # pure constant (+ white): any kind with c1 = 0
            kind, ls, c1 = KIND_RBF, (1.0,) * dim, 0.0
        return KernelSpec(kind, float(c0), float(c1), kdiag, np.asarray(ls, dtype=np.float64))
    remap = {old: new for new, old in enumerate(used)}
    nf, nt = len(used), len(terms)
    texp = np.zeros((nt, nf), dtype=np.int32)
    for t, (_, idx) in enumerate(terms):
        for i in idx:
            texp[t, remap[i]] += 1
    if nf > GEN_MAX_F or nt > GEN_MAX_T or texp.max() > GEN_MAX_EXP:
        raise NotImplementedError(
            "kernel %r expands to %d stationary leaves and %d product terms; the device path supports up to "
            "%d leaves, %d terms and powers <= %d" % (kernel, nf, nt, GEN_MAX_F, GEN_MAX_T, GEN_MAX_EXP))
    padded = next((P for P in (2, 4, 6, 8, 10, 12, 16, 20, 24, 32) if dim <= P), dim)     # csrc/model.h: padded_dim
    if nf * padded > GEN_MAX_FP:
        raise NotImplementedError(
            "kernel %r: %d stationary leaves at input dimension %d exceed the staging buffers of the general "
            "device kernels (leaves x padded dimension <= %d)" % (kernel, nf, dim, GEN_MAX_FP))
    general = GeneralSpec([leaves[i][0] for i in used], [leaves[i][1] for i in used],
                          [c for c, _ in terms], texp)
    return KernelSpec(KIND_GENERAL, 0.0, 0.0, kdiag, np.ones(dim), general)
