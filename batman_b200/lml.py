"""Batched log-marginal-likelihood on the device (sklearn:_gpr.py:584-617).

For R candidate hyper-parameter vectors at once:
    K_r = k_theta_r(X, X) + 1e-10 I ;  L_r = chol(K_r) ;  alpha_r = K_r^-1 y
    LML_r = -1/2 y.alpha_r - sum(log diag L_r) - n/2 log(2 pi)     (-inf if chol fails)
"""
import math

import numpy as np
import torch

from . import kernel_spec as ks

JITTER = 1e-10          # GaussianProcessRegressor(alpha=1e-10), sklearn:_gpr.py:215
_MAX_BATCH_BYTES = 8 << 30


def _stat(kind, s):
    """Same expressions as csrc/covariance.cuh, on the squared scaled distance."""
    if kind == ks.KIND_RBF:
        return torch.exp(-0.5 * s)
    d = torch.sqrt(s)
    if kind == ks.KIND_MATERN12:
        return torch.exp(-d)
    if kind == ks.KIND_MATERN32:
        t = d * math.sqrt(3)
        return (1.0 + t) * torch.exp(-t)
    if kind == ks.KIND_MATERN52:
        t = d * math.sqrt(5)
        return (1.0 + t + t * t / 3.0) * torch.exp(-t)
    return torch.exp(-(d * d) / 2.0)


def build_K(X_dev, specs):
    """(R, n, n) training covariance incl. jitter; diagonal = kernel_.diag + 1e-10 exactly."""
    n, p = X_dev.shape
    R = len(specs)
    K = torch.empty((R, n, n), dtype=torch.float64, device=X_dev.device)
    eye = torch.eye(n, dtype=torch.bool, device=X_dev.device)
    for r, spec in enumerate(specs):
        ls = torch.from_numpy(np.asarray(spec.ls, dtype=np.float64)).to(X_dev.device)
        U = X_dev / ls
        s = torch.zeros((n, n), dtype=torch.float64, device=X_dev.device)
        for c in range(p):                       # cdist order: sequential, non-fused
            d = U[:, c, None] - U[None, :, c]
            s = s + d * d
        Kr = spec.c0 + spec.c1 * _stat(spec.kind, s)
        Kr[eye] = spec.kdiag + JITTER
        K[r] = Kr
    return K


def lml_batched(X_dev, y_dev, specs):
    """-> numpy (R,) of LML values; -inf where K is not numerically PD."""
    n = X_dev.shape[0]
    out = np.empty(len(specs))
    per = max(1, int(_MAX_BATCH_BYTES // (3 * n * n * 8)))
    for lo in range(0, len(specs), per):
        K = build_K(X_dev, specs[lo:lo + per])
        L, info = torch.linalg.cholesky_ex(K)
        alpha = torch.cholesky_solve(y_dev[None, :, None].expand(K.shape[0], n, 1), L)[..., 0]
        val = -0.5 * (alpha @ y_dev) - torch.log(torch.diagonal(L, dim1=1, dim2=2)).sum(dim=1) \
            - n / 2 * math.log(2 * math.pi)
        val = torch.where(info == 0, val, torch.full_like(val, -float('inf')))
        out[lo:lo + per] = val.cpu().numpy()
    return out


def factorise(X_dev, y_dev, spec):
    """Final state at the chosen theta (sklearn:_gpr.py:349-367): L, alpha on the host + LML."""
    K = build_K(X_dev, [spec])[0]
    L, info = torch.linalg.cholesky_ex(K)
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError(
            "The kernel is not returning a positive definite matrix. Try gradually increasing the "
            "'alpha' parameter of your GaussianProcessRegressor estimator.")   # sklearn:_gpr.py:354-360
    alpha = torch.cholesky_solve(y_dev[:, None], L)[:, 0]
    n = X_dev.shape[0]
    val = -0.5 * float(alpha @ y_dev) - float(torch.log(torch.diagonal(L)).sum()) - n / 2 * math.log(2 * math.pi)
    return torch.tril(L).cpu().numpy(), alpha.cpu().numpy(), val
