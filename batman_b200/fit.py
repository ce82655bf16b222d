"""Hyper-parameter fit of one GP output on the device.

Mirrors what ``Kriging.__init__`` asks scikit-learn to do (kriging.py:99-123,
151-190; sklearn:_gpr.py:302-367, 541-617):

* objective        -LML(theta), theta = log hyper-parameters, y normalised by
                   mean AND std (sklearn:_gpr.py:275-280);
* global optimiser 3 differential-evolution runs, popsize 15+i, tol 1e-3
                   (kriging.py:170-171), best of three (kriging.py:183-190).
                   The reference forks one process per run and evaluates one
                   theta at a time (and pays for a gradient it discards,
                   kriging.py:164-166); here scipy's DE is driven in its
                   ``vectorized``/``deferred`` mode so that each generation is
                   ONE batched LML call on the GPU.
* local optimiser  L-BFGS-B from theta0 plus ``10*dim`` uniform restarts in
                   log-bounds (kriging.py:105; sklearn:_gpr.py:311-330), with
                   central finite differences evaluated as one batch.
* final state      K = k_theta(X, X) + 1e-10 I, L = chol(K), alpha = K^-1 y
                   (sklearn:_gpr.py:349-367).

The batched LML itself lives in ``lml.py`` (device kernels through the C ABI).
Parity for fit is defined on LML(theta) values at given theta and on
predictions at a given theta -- the DE path is unseeded in the reference, so
its theta is not reproducible even against itself.
"""
import numpy as np
from scipy.optimize import differential_evolution, minimize

from . import kernel_spec
from .lml import lml_batched, factorise


def normalise_y(y):
    """sklearn:_gpr.py:275-280 (``_handle_zeros_in_scale`` included)."""
    y = np.asarray(y, dtype=np.float64)
    mean = float(np.mean(y, axis=0))
    std = float(np.std(y, axis=0))
    if std < 10 * np.finfo(np.float64).eps:
        std = 1.0
    return (y - mean) / std, mean, std


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _shared_seed(device):
    """One integer drawn on rank 0 and broadcast: the (unseeded, kriging.py:170-171) DE runs must walk the
    same trajectory on every rank for the sharded objective to make sense."""
    import torch
    import torch.distributed as dist
    world, _ = _world()
    seed = torch.tensor([np.random.randint(0, 2 ** 31 - 1)], dtype=torch.int64, device=device)
    if world > 1:
        dist.broadcast(seed, src=0)
    return int(seed.item())


class _Objective:
    def __init__(self, kernel, X_dev, y_dev, dim):
        self.kernel, self.X_dev, self.y_dev, self.dim = kernel, X_dev, y_dev, dim
        self.n_evals = 0

    def specs(self, thetas):
        return [kernel_spec.flatten(self.kernel.clone_with_theta(t), self.dim) for t in thetas]

    def neg_lml(self, thetas):
        """thetas (R, n_theta) -> -LML (R,), +inf where the Cholesky fails (sklearn:_gpr.py:591-593).

        With torch.distributed initialised the R candidates (a DE generation, or the finite-difference stencil
        of an L-BFGS-B restart) are sharded over the ranks -- each GPU factorises its slice -- and the values
        are all-gathered, so every rank sees the same objective and the optimisers stay in lockstep."""
        thetas = np.atleast_2d(thetas)
        self.n_evals += len(thetas)
        world, rank = _world()
        if world == 1 or len(thetas) < world:
            return -lml_batched(self.X_dev, self.y_dev, self.specs(thetas))
        import torch
        import torch.distributed as dist
        lo, hi = rank * len(thetas) // world, (rank + 1) * len(thetas) // world
        mine = -lml_batched(self.X_dev, self.y_dev, self.specs(thetas[lo:hi]))
        per = (len(thetas) + world - 1) // world
        buf = torch.full((per,), float('nan'), dtype=torch.float64, device=self.X_dev.device)
        buf[:hi - lo] = torch.from_numpy(mine).to(self.X_dev.device)
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf)
        out = np.empty(len(thetas))
        for r in range(world):
            rlo, rhi = r * len(thetas) // world, (r + 1) * len(thetas) // world
            out[rlo:rhi] = gathered[r][:rhi - rlo].cpu().numpy()
        return out


def optimise_theta(kernel, X_dev, y_dev, dim, global_optimizer, n_restarts_local):
    theta0 = np.asarray(kernel.theta, dtype=np.float64)
    if theta0.size == 0:
        return theta0, None
    bounds = np.asarray(kernel.bounds, dtype=np.float64)
    obj = _Objective(kernel, X_dev, y_dev, dim)
    if global_optimizer:
        best = None
        seed = _shared_seed(X_dev.device)
        for i in range(3):                                           # kriging.py:176-177 (n_restart = 3)
            res = differential_evolution(lambda th: obj.neg_lml(th.T), bounds, tol=0.001, popsize=15 + i,
                                         vectorized=True, updating='deferred', polish=False, seed=seed + i)
            # scipy's default polish (L-BFGS-B on the best member) with batched central differences
            res = _lbfgs(obj, res.x, bounds, fallback=res)
            if best is None or res.fun < best.fun:
                best = res
        return best.x, -best.fun
    starts = [theta0]
    if n_restarts_local > 0:
        if not np.isfinite(bounds).all():
            raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all "
                             "bounds are finite.")                    # sklearn:_gpr.py:313-317
        rng = np.random.RandomState(_shared_seed(X_dev.device)) if _world()[0] > 1 else np.random
        starts += [rng.uniform(bounds[:, 0], bounds[:, 1]) for _ in range(n_restarts_local)]
    best = None
    for s in starts:
        res = _lbfgs(obj, s, bounds)
        if best is None or res.fun < best.fun:
            best = res
    return best.x, -best.fun


def _lbfgs(obj, x0, bounds, fallback=None, h=1e-6):
    d = len(x0)

    def fun_and_grad(x):
        pts = np.repeat(x[None, :], 2 * d + 1, axis=0)
        for c in range(d):
            pts[1 + 2 * c, c] = min(x[c] + h, bounds[c, 1])
            pts[2 + 2 * c, c] = max(x[c] - h, bounds[c, 0])
        vals = obj.neg_lml(pts)
        if not np.isfinite(vals[0]):
            return 1e25, np.zeros(d)
        g = np.zeros(d)
        for c in range(d):
            dx = pts[1 + 2 * c, c] - pts[2 + 2 * c, c]
            if dx > 0 and np.isfinite(vals[1 + 2 * c]) and np.isfinite(vals[2 + 2 * c]):
                g[c] = (vals[1 + 2 * c] - vals[2 + 2 * c]) / dx
        return float(vals[0]), g

    res = minimize(fun_and_grad, x0, jac=True, method='L-BFGS-B', bounds=bounds)
    if fallback is not None and not (res.fun < fallback.fun):
        return fallback
    return res


def fit_output(kernel, X_dev, X_host, y, global_optimizer, n_restarts_local):
    """-> (kernel_, spec, alpha (n,), L (n, n), y_mean, y_std, lml_value, packed L^-1 tiles on the device)."""
    dim = X_host.shape[1]
    yn, y_mean, y_std = normalise_y(y)
    import torch
    y_dev = torch.from_numpy(np.ascontiguousarray(yn)).to(X_dev.device)
    theta, lml_val = optimise_theta(kernel, X_dev, y_dev, dim, global_optimizer, n_restarts_local)
    kernel_ = kernel.clone_with_theta(theta)
    spec = kernel_spec.flatten(kernel_, dim)
    L, alpha, lml_final, wt = factorise(X_dev, y_dev, spec)
    return kernel_, spec, alpha, L, y_mean, y_std, lml_final, wt
