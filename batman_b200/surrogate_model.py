"""``SurrogateModel`` facade for ``kind='kriging'`` (surrogate_model.py:46-198 of the reference).

Only the part of the facade that sits on the hot path is mirrored: unit-cube
scaling with a MinMaxScaler fitted on the corners (surrogate_model.py:108-109,
139, 181), dispatch to ``Kriging.evaluate`` (:184) and the POD inverse
transform applied to the mean AND to sigma (:191-196 -- statistically odd, but
it is the reference behaviour and parity reproduces it).  Other surrogate kinds
are out of scope and raise.
"""
import logging

import numpy as np

from .kriging import Kriging


class MinMax:
    """sklearn ``MinMaxScaler().fit(corners)`` restated: transform is ``X * scale_ + min_``."""

    def __init__(self, corners):
        corners = np.asarray(corners, dtype=np.float64)
        data_min, data_max = corners.min(axis=0), corners.max(axis=0)
        data_range = data_max - data_min
        data_range[data_range == 0.0] = 1.0          # _handle_zeros_in_scale
        self.scale_ = 1.0 / data_range
        self.min_ = 0.0 - data_min * self.scale_

    def transform(self, X):
        X = np.array(X, dtype=np.float64, copy=True)
        X *= self.scale_
        X += self.min_
        return X


class Pod:
    """Holder of a fitted POD basis for the inverse transform (pod/pod.py:219-228).

    Duck-compatible with ``batman.pod.Pod`` for this path: ``mean_snapshot`` (F,),
    ``U`` (F, m) and ``inverse_transform``.  The SVD fit itself (pod.py:198-217)
    is a one-off n_snapshots x n_features problem and stays with the reference.
    """

    def __init__(self, mean_snapshot, U):
        self.mean_snapshot = np.ascontiguousarray(mean_snapshot, dtype=np.float64)
        self.U = np.ascontiguousarray(U, dtype=np.float64)
        self._dev = None

    def device_basis(self, device):
        import torch
        if self._dev is None or self._dev[0].device != device:
            self._dev = (torch.from_numpy(self.mean_snapshot).to(device), torch.from_numpy(self.U).to(device))
        return self._dev

    def inverse_transform_device(self, modes_dev):
        """(k, m) CUDA tensor -> (k, F): mean_snapshot + (U @ modes.T).T"""
        from .pod_kernels import pod_inverse
        mean_d, U_d = self.device_basis(modes_dev.device)
        return pod_inverse(mean_d, U_d, modes_dev)

    def inverse_transform(self, samples):
        import torch
        samples = np.atleast_2d(np.asarray(samples, dtype=np.float64))
        out = self.inverse_transform_device(torch.from_numpy(np.ascontiguousarray(samples)).cuda())
        return np.atleast_2d(out.cpu().numpy())

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_dev'] = None
        return state


class _SpaceInfo(list):
    """The two attributes of ``batman.space.Space`` this path reads (uq.py:174, surrogate_model.py:160-162)."""
    doe_init = 0


class SurrogateModel:
    """Surrogate model (kriging only)."""

    logger = logging.getLogger(__name__)

    def __init__(self, kind, corners, plabels=None, **kwargs):
        if kind != 'kriging':
            raise NotImplementedError("batman_b200 accelerates kind='kriging' only (got %r); use the "
                                      "reference SurrogateModel for other kinds" % (kind,))
        self.kind = kind
        self.scaler = MinMax(corners)
        self.corners = np.asarray(corners, dtype=np.float64)
        self.plabels = plabels
        self.space = _SpaceInfo()
        self.data = None
        self.pod = None
        self.update = False
        self.settings = kwargs
        self.predictor = None

    def fit(self, sample, data, pod=None):
        """surrogate_model.py:128-164."""
        self.data = data
        sample = np.array(sample, dtype=np.float64)
        sample_scaled = self.scaler.transform(sample)
        self.logger.info('Creating predictor of kind {}...'.format(self.kind))
        self.predictor = Kriging(sample_scaled, np.asarray(data, dtype=np.float64), **self.settings)
        self._adopt(pod, sample)
        self.logger.info('Predictor created')

    def adopt(self, predictor, sample, pod=None):
        """Attach an already fitted predictor (e.g. ``Kriging.from_state`` on a reference pickle)."""
        self.predictor = predictor
        self._adopt(pod, np.asarray(sample, dtype=np.float64))

    def _adopt(self, pod, sample):
        if pod is not None and not isinstance(pod, Pod):
            pod = Pod(pod.mean_snapshot, pod.U)      # batman.pod.Pod or any duck
        self.pod = pod
        del self.space[:]
        self.space.extend(np.asarray(sample).tolist())
        self.space.doe_init = len(sample)
        self.update = False

    def estimate_quality(self, method='LOO'):
        """Leave-one-out Q2 and the point of largest error (surrogate_model.py:200-269).

        As in the reference every fold REFITS the predictor (hyper-parameters included) on n - 1 points and predicts
        the held-out one; the folds run one after the other on the device instead of in a fork pool."""
        from sklearn.metrics import r2_score
        if self.pod is not None:
            raise NotImplementedError("estimate_quality with a POD is Pod.estimate_quality (pod.py:264-420): "
                                      "use the reference for that path")
        self.logger.info('Estimating Surrogate quality...')
        sample = np.array(self.space, dtype=np.float64)
        data = np.asarray(self.data, dtype=np.float64).reshape(len(sample), -1)
        sample_scaled = self.scaler.transform(sample)
        y_pred = np.empty_like(data)
        keep = np.ones(len(sample), dtype=bool)
        for i in range(len(sample)):
            keep[i] = False
            fold = Kriging(sample_scaled[keep], data[keep], **self.settings)
            y_pred[i] = fold.evaluate(sample_scaled[i])[0][0]
            keep[i] = True
        q2_loo = r2_score(data, y_pred)
        index = ((data - y_pred) ** 2).sum(axis=1).argmax()
        point = self.space[index]
        self.logger.info('Surrogate quality: {}, max error location at {}'.format(q2_loo, point))
        return q2_loo, point

    # ------------------------------------------------------------------ device-level call used by UQ
    def predict_device(self, raw_points_dev, return_std=True):
        """raw (unscaled) points on the device -> (results, sigma) on the device, POD applied."""
        with self.predictor.input_scaler(self.scaler.scale_, self.scaler.min_) as model:
            mean, sigma = model.predict_device(raw_points_dev, return_std)
        if self.pod is not None:
            mean = self.pod.inverse_transform_device(mean)
            if sigma is not None:
                sigma = self.pod.inverse_transform_device(sigma)     # surrogate_model.py:193-196
        return mean, sigma

    def __call__(self, points):
        """Predict snapshots (surrogate_model.py:166-198): returns (results, sigma), each (k, F)."""
        import torch
        points = np.atleast_2d(np.asarray(points, dtype=np.float64))
        n_features = self.corners.shape[1]
        if points.shape[1] != n_features:
            raise ValueError("X has %d features, but MinMaxScaler is expecting %d features as input."
                             % (points.shape[1], n_features))
        if self.pod is None:
            # numpy in / numpy out with the copies pipelined against the kernels (DeviceModel.predict)
            with self.predictor.input_scaler(self.scaler.scale_, self.scaler.min_) as model:
                mean, sigma = model.predict(points, return_std=True)
            return np.atleast_2d(mean), sigma
        pts = torch.from_numpy(np.ascontiguousarray(points)).cuda()
        mean, sigma = self.predict_device(pts, return_std=True)
        return np.atleast_2d(mean.cpu().numpy()), sigma.cpu().numpy()
