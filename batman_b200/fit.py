"""Hyper-parameter fit of one GP output on the device.

Mirrors what ``Kriging.__init__`` asks scikit-learn to do (kriging.py:99-123,
151-190; sklearn:_gpr.py:302-367, 541-617):

* objective        -LML(theta), theta = log hyper-parameters, y normalised by
                   mean AND std (sklearn:_gpr.py:275-280);
* global optimiser ``n_restart`` (3) differential-evolution runs, popsize
                   15+i, tol 1e-3 (kriging.py:170-171), best of them
                   (kriging.py:183-190).  The reference forks one process per
                   run and evaluates one theta at a time (and pays for a
                   gradient it discards, kriging.py:164-166); here every run of
                   every output is a thread driving scipy's DE in its
                   ``vectorized``/``deferred`` mode, and the generations of ALL
                   runs are ONE batched LML call on the GPU (LockstepBatcher).
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
import threading

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


class LockstepBatcher:
    """Rendezvous of optimiser threads: ONE device batch per generation for all of them.

    The reference forks one process per differential-evolution restart (kriging.py:109-115, 176-181) and fits the
    outputs one after the other (kriging.py:117-143); every process evaluates one theta at a time.  Here every
    (output, restart) pair is a client thread running an unmodified scipy optimiser in its vectorised mode; a call
    to the objective parks the thread, and when ALL live clients are parked their candidates are concatenated (in
    client order) into one ``lml_batched`` call.  A batch fires only when every live client waits or has retired, so
    its composition does not depend on thread timing -- with ``torch.distributed`` every rank builds the same batches
    and the candidates can be sharded over the ranks."""

    def __init__(self, evaluate, clients):
        self._evaluate = evaluate
        self._cond = threading.Condition()
        self._active = set(clients)
        self._pending, self._results = {}, {}
        self._error = None
        self.n_batches = self.n_evals = self.largest_batch = 0

    def _fire_if_ready(self):                                   # lock held
        if not self._pending or set(self._pending) != self._active:
            return
        batch = sorted(self._pending.items())
        self._pending = {}
        try:
            values = self._evaluate(batch)
        except BaseException as exc:                            # every parked client must wake up and fail
            self._error = exc
        else:
            for (cid, _), v in zip(batch, values):
                self._results[cid] = v
            n = sum(len(t) for _, t in batch)
            self.n_batches += 1
            self.n_evals += n
            self.largest_batch = max(self.largest_batch, n)
        self._cond.notify_all()

    def submit(self, cid, thetas):
        with self._cond:
            self._pending[cid] = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
            self._fire_if_ready()
            while cid not in self._results and self._error is None:
                self._cond.wait()
            if self._error is not None:
                raise RuntimeError("batched LML evaluation failed") from self._error
            return self._results.pop(cid)

    def retire(self, cid):
        with self._cond:
            self._active.discard(cid)
            self._pending.pop(cid, None)
            self._fire_if_ready()


class _Objective:
    """-LML of (output q, theta) pairs, evaluated for a whole lockstep batch at once."""

    def __init__(self, kernel, X_dev, Y_dev, dim):
        self.kernel, self.X_dev, self.Y_dev, self.dim = kernel, X_dev, Y_dev, dim

    def specs(self, thetas):
        return [kernel_spec.flatten(self.kernel.clone_with_theta(t), self.dim) for t in thetas]

    def neg_lml(self, thetas, targets):
        """thetas (R, n_theta), targets (R,) -> -LML (R,), +inf where the Cholesky fails (sklearn:_gpr.py:591-593).

        With torch.distributed initialised the R candidates are sharded over the ranks -- each GPU factorises its
        slice -- and the values are all-gathered, so every rank sees the same objective and the optimisers stay in
        lockstep."""
        thetas = np.atleast_2d(thetas)
        targets = np.asarray(targets, dtype=np.int64)
        world, rank = _world()
        if world == 1 or len(thetas) < world:
            return -lml_batched(self.X_dev, self.Y_dev, self.specs(thetas), targets)
        import torch
        import torch.distributed as dist
        lo, hi = rank * len(thetas) // world, (rank + 1) * len(thetas) // world
        mine = -lml_batched(self.X_dev, self.Y_dev, self.specs(thetas[lo:hi]), targets[lo:hi])
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

    def evaluate_batch(self, client_target):
        """-> the ``evaluate`` callback of a LockstepBatcher whose client ``cid`` scores output ``client_target[cid]``."""
        def evaluate(batch):
            thetas = np.concatenate([t for _, t in batch], axis=0)
            targets = np.concatenate([np.full(len(t), client_target[cid]) for cid, t in batch])
            values = self.neg_lml(thetas, targets)
            out, o = [], 0
            for _, t in batch:
                out.append(values[o:o + len(t)])
                o += len(t)
            return out
        return evaluate


MAX_CLIENTS = 256        # optimiser threads alive at once (they are parked almost all the time)


def optimise_thetas(kernel, X_dev, Y_dev, dim, global_optimizer, n_restarts_local, n_restart=3):
    """Hyper-parameters of ALL outputs: -> (thetas (m, n_theta), lml values (m,) or None, stats dict).

    global_optimizer: ``n_restart`` differential-evolution runs per output, popsize 15 + i, tol 1e-3, best of them
    (kriging.py:170-171, 183-190), each followed by scipy's default polish (L-BFGS-B on the best member, here with
    batched central differences).  Otherwise L-BFGS-B from theta0 plus ``n_restarts_local`` uniform restarts in
    log-bounds (kriging.py:105; sklearn:_gpr.py:311-330).  Every run of every output is a client of one
    LockstepBatcher."""
    import torch
    m = Y_dev.shape[0]
    theta0 = np.asarray(kernel.theta, dtype=np.float64)
    if theta0.size == 0:
        return np.tile(theta0, (m, 1)), None, dict(batches=0, evaluations=0, largest_batch=0, clients=0)
    bounds = np.asarray(kernel.bounds, dtype=np.float64)
    obj = _Objective(kernel, X_dev, Y_dev, dim)
    seed = _shared_seed(X_dev.device)
    if global_optimizer:
        runs = [('de', i) for i in range(n_restart)]
    else:
        if n_restarts_local > 0 and not np.isfinite(bounds).all():
            raise ValueError("Multiple optimizer restarts (n_restarts_optimizer>0) requires that all "
                             "bounds are finite.")                    # sklearn:_gpr.py:313-317
        runs = [('lbfgs', i) for i in range(1 + n_restarts_local)]
    per_group = max(1, MAX_CLIENTS // len(runs))
    best = [None] * m
    stats = dict(batches=0, evaluations=0, largest_batch=0, clients=0)
    device = X_dev.device
    for q0 in range(0, m, per_group):
        clients = [(q, r) for q in range(q0, min(m, q0 + per_group)) for r in range(len(runs))]
        batcher = LockstepBatcher(obj.evaluate_batch({c: c[0] for c in clients}), clients)
        results, errors = {}, []

        def work(cid):
            q, r = cid
            kind, i = runs[r]
            try:
                torch.cuda.set_device(device)                    # the current device is per thread
                fun = lambda pts: batcher.submit(cid, pts)       # noqa: E731
                if kind == 'de':
                    res = differential_evolution(lambda th: fun(th.T), bounds, tol=0.001, popsize=15 + i,
                                                 vectorized=True, updating='deferred', polish=False,
                                                 seed=seed + 7919 * q + i)
                    res = _lbfgs(fun, res.x, bounds, fallback=res)
                else:
                    if i == 0:
                        x0 = theta0
                    else:
                        rng = np.random.RandomState((seed + 7919 * q + i) % (2 ** 31 - 1))
                        x0 = rng.uniform(bounds[:, 0], bounds[:, 1])
                    res = _lbfgs(fun, x0, bounds)
                results[cid] = res
            except BaseException as exc:
                errors.append(exc)
            finally:
                batcher.retire(cid)

        threads = [threading.Thread(target=work, args=(c,), daemon=True) for c in clients]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        for (q, r), res in sorted(results.items()):
            if best[q] is None or res.fun < best[q].fun:
                best[q] = res
        stats['batches'] += batcher.n_batches
        stats['evaluations'] += batcher.n_evals
        stats['largest_batch'] = max(stats['largest_batch'], batcher.largest_batch)
        stats['clients'] = max(stats['clients'], len(clients))
    return np.array([b.x for b in best]), np.array([-b.fun for b in best]), stats


def _lbfgs(neg_lml, x0, bounds, fallback=None, h=1e-6):
    d = len(x0)

    def fun_and_grad(x):
        pts = np.repeat(x[None, :], 2 * d + 1, axis=0)
        for c in range(d):
            pts[1 + 2 * c, c] = min(x[c] + h, bounds[c, 1])
            pts[2 + 2 * c, c] = max(x[c] - h, bounds[c, 0])
        vals = neg_lml(pts)
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


def fit_outputs(kernel, X_dev, X_host, data, global_optimizer, n_restarts_local, n_restart=3):
    """All outputs of one Kriging model -> ([(kernel_, spec, alpha (n,), L (n, n), y_mean, y_std, lml_value,
    packed L^-1 tiles on the device) per output], optimiser statistics)."""
    import torch
    dim = X_host.shape[1]
    data = np.asarray(data, dtype=np.float64)
    norm = [normalise_y(data[:, q]) for q in range(data.shape[1])]
    Y_dev = torch.from_numpy(np.ascontiguousarray(np.stack([nq[0] for nq in norm]))).to(X_dev.device)
    thetas, _, stats = optimise_thetas(kernel, X_dev, Y_dev, dim, global_optimizer, n_restarts_local, n_restart)
    fitted = []
    for q, (yn, y_mean, y_std) in enumerate(norm):
        kernel_ = kernel.clone_with_theta(thetas[q])
        spec = kernel_spec.flatten(kernel_, dim)
        L, alpha, lml_final, wt = factorise(X_dev, Y_dev[q], spec)
        fitted.append((kernel_, spec, alpha, L, y_mean, y_std, lml_final, wt))
    return fitted, stats


def fit_output(kernel, X_dev, X_host, y, global_optimizer, n_restarts_local, n_restart=3):
    """One output (kept for callers that fit a single column)."""
    fitted, _ = fit_outputs(kernel, X_dev, X_host, np.asarray(y, dtype=np.float64).reshape(-1, 1), global_optimizer,
                            n_restarts_local, n_restart)
    return fitted[0]
