"""Resampling criteria that query sigma(x) (space/refiner.py:287-325, 523-592), batched.

The reference maximises each criterion with ``scipy.optimize.differential_evolution``
calling ``surrogate(x)`` one point at a time (misc.py:239-264).  Here the same
optimiser runs in its ``vectorized`` mode: every generation is ONE batched
``SurrogateModel.__call__`` on the device.  Continuous parameters only (the
reference's per-value loop over a discrete parameter is host logic above this).
"""
import logging
import warnings

import numpy as np
from scipy.optimize import differential_evolution
from scipy.stats import norm


class Refiner:
    """Sigma / expected-improvement point selection on a fitted ``batman_b200.SurrogateModel``."""

    logger = logging.getLogger(__name__)

    def __init__(self, surrogate, corners, delta_space=0.08, pod=None):
        """Arguments as refiner.py:44-95 (``data`` must be a fitted SurrogateModel)."""
        self.surrogate = surrogate
        self.pod_S = 1 if pod is None else np.asarray(pod.S, dtype=np.float64)
        self.points = np.array(surrogate.space[:], dtype=np.float64)
        self.corners = np.array(corners, dtype=np.float64).T                 # (dim, 2)
        self.dim = len(self.corners)
        for i in range(self.dim):                                            # inner contraction, refiner.py:86-90
            lo, hi = self.corners[i]
            self.corners[i, 0] = lo + delta_space * (hi - lo)
            self.corners[i, 1] = hi - delta_space * (self.corners[i, 1] - self.corners[i, 0])
        self._lo, self._hi = self.corners[:, 0].copy(), self.corners[:, 1].copy()
        self.points_scaled = (self.points - self._lo) / (self._hi - self._lo)   # MinMaxScaler on the corners

    # -- weighted prediction / sigma (refiner.py:97-130), batched over points
    def _weights(self, n_out):
        w = np.broadcast_to(np.asarray(self.pod_S, dtype=np.float64) ** 2, (n_out,)) if np.ndim(self.pod_S) else \
            np.ones(n_out)
        return w / w.sum()

    def pred_sigma(self, coords):
        f, sigma = self.surrogate(np.atleast_2d(coords))
        w = self._weights(f.shape[1])
        return f @ w, sigma @ w

    def _minimise(self, fun):
        with np.errstate(divide='ignore', invalid='ignore'):
            res = differential_evolution(lambda x: fun(x.T), self.corners, vectorized=True, updating='deferred')
        return res.x, res.fun

    def sigma(self):
        """Point of maximum ``sum S_i^2 sigma_i`` (refiner.py:287-325)."""
        s2 = np.asarray(self.pod_S, dtype=np.float64) ** 2

        def func_sigma(x):
            _, sigma = self.surrogate(x)
            return -np.sum(s2 * sigma, axis=1)
        return self._minimise(func_sigma)[0]

    def optimization(self, method='EI', extremum='min'):
        """Maximisation of the probability / expected improvement (refiner.py:523-592)."""
        sign = 1 if extremum == 'min' else -1
        gen = sign * self.pred_sigma(self.points)[0]
        min_value = float(gen.min())
        target = min_value - 0.1 * np.abs(min_value)

        def criterion(x):
            xs = (x - self._lo) / (self._hi - self._lo)
            # refiner.py:548-552 takes np.linalg.norm(x_scaled[0][no_discrete] - p[no_discrete], -1) where the index
            # (None, or a list of lists) makes the operand a 1 x dim MATRIX: ord=-1 is then the smallest column sum
            # of absolute values, i.e. min_i |dx_i| (not the vector (-1)-norm (sum |dx_i|^-1)^-1).
            d = np.abs(xs[:, None, :] - self.points_scaled[None, :, :]).min(axis=2)
            too_close = (d < 0.02).any(axis=1)
            pred, sigma = self.pred_sigma(x)
            pred = sign * pred
            std_dev = np.sqrt(sigma)
            if method == 'PI':
                val = -norm.cdf((target - pred) / std_dev)
            else:
                diff = min_value - pred
                val = -(diff * norm.cdf(diff / std_dev) + std_dev * norm.pdf(diff / std_dev))
            return np.where(too_close, np.inf, val)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return self._minimise(criterion)[0]
