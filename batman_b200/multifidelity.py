"""Evofusion multifidelity predictor on the device (surrogate/multifidelity.py:23-80).

Two Kriging models -- the cheap function, and the error between the expensive
function and the cheap one at the shared points -- summed at evaluation
(multifidelity.py:75-78).  Both fits and both evaluations run through
``batman_b200.Kriging`` (batched, no per-point ``multi_eval`` loop).
"""
import logging

import numpy as np

from .kriging import Kriging


class Evofusion:
    """Multifidelity algorithm using Evofusion."""

    logger = logging.getLogger(__name__)

    def __init__(self, sample, data, **kriging_settings):
        """``sample``: (n_e + n_c, 1 + n_features), first column = fidelity flag (0 = expensive, 1 = cheap),
        expensive rows first (``Space(..., multifidelity=...)`` layout, space.py:163-166);
        ``data``: (n_e + n_c, [n_outputs]) in the same order (multifidelity.py:29-50)."""
        sample = np.array(sample, dtype=np.float64)
        data = np.array(data, dtype=np.float64)
        sample = [sample[sample[:, 0] == 0][:, 1:], sample[sample[:, 0] == 1][:, 1:]]
        n_e, n_c = sample[0].shape[0], sample[1].shape[0]
        data = [data[:n_e].reshape((n_e, -1)), data[n_e:].reshape((n_c, -1))]
        self.model_c = Kriging(sample[1], data[1], **kriging_settings)          # low fidelity model
        # cheap twin of every expensive point.  The reference's index arithmetic (multifidelity.py:55-56)
        # relies on the expensive points being the first n_e cheap points; here the coordinates are matched.
        twin = np.full(n_e, -1)
        for i, x in enumerate(sample[0]):
            hit = np.flatnonzero(np.all(sample[1] == x, axis=1))
            if hit.size:
                twin[i] = hit[0]
        if np.any(twin < 0):
            raise ValueError("Evofusion: every expensive point needs a cheap evaluation at the same coordinates")
        data_err = data[0] - data[1][twin]
        self.model_err = Kriging(sample[0], data_err, **kriging_settings)       # error model between fidelities

    def evaluate(self, point):
        """(n_points, n_features) -> (prediction, sigma), each (n_points, n_outputs) (multifidelity.py:63-80)."""
        f_c, sigma_c = self.model_c.evaluate(point)
        f_err, sigma_err = self.model_err.evaluate(point)
        return f_c + f_err, sigma_c + sigma_err
