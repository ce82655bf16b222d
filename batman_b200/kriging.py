# coding: utf8
"""
Kriging Class (B200 drop-in)
============================

Same constructor, attributes and ``evaluate`` contract as
``batman.surrogate.Kriging`` (batman/surrogate/kriging.py:41-214)::

    >> from batman_b200 import Kriging
    >> import numpy as np
    >> sample = np.array([[2, 4], [3, 5], [6, 9]])
    >> data = np.array([[12, 1], [10, 2], [9, 4]])
    >> predictor = Kriging(sample, data)
    >> predictor.evaluate((5.0, 8.0))
    (array([[10.333,  3.591]]), array([[1.247, 0.694]]))

One independent GP per output column (kriging.py:117-148).  The per-point
python loop of ``multi_eval`` (functions/utils.py:19-31) + ``gp.predict`` is
replaced by one batched device call; shapes are the ones ``multi_eval``
produces: ``(n_points, n_outputs)`` for both mean and sigma, a single 1-D point
giving ``(1, n_outputs)``.
"""
import contextlib
import logging

import numpy as np
from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

from . import kernel_spec


class FittedGP:
    """What ``Kriging.gp_models[i]`` exposes of a fitted GaussianProcessRegressor
    (sklearn:_gpr.py:349-367): plain numpy state, picklable."""

    def __init__(self, kernel_, X_train_, alpha_, L_, y_train_mean, y_train_std, lml=None):
        self.kernel_ = kernel_
        self.X_train_ = X_train_
        self.alpha_ = alpha_
        self.L_ = L_
        self._y_train_mean = y_train_mean
        self._y_train_std = y_train_std
        self.log_marginal_likelihood_value_ = lml


class Kriging:
    """Kriging based on Gaussian Process, evaluated on a B200."""

    logger = logging.getLogger(__name__)

    def __init__(self, sample, data, kernel=None, noise=False, global_optimizer=True, n_restart=None):
        r"""Create the predictor (arguments as kriging.py:41-76).

        :param array_like sample: Sample used to generate the data
          (n_samples, n_features), already scaled to the unit cube.
        :param array_like data: Observed data (n_samples, n_outputs), 2-D.
        :param kernel: Kernel from scikit-learn
          (:class:`sklearn.gaussian_process.kernels`.*).
        :param float/bool noise: Noise used into kriging.
        :param bool global_optimizer: Whether to do global optimization or
          gradient based optimization to estimate hyperparameters.
        :param int n_restart: (addition) number of differential-evolution
          restarts; the reference hard-wires 3 (kriging.py:101).  All restarts
          of all outputs advance in lockstep, one batched device call per
          generation.
        """
        sample = np.atleast_2d(np.asarray(sample, dtype=np.float64))
        data = np.asarray(data, dtype=np.float64)
        dim = sample.shape[1]
        self.model_len = data.shape[1]                                   # kriging.py:78
        if kernel is not None:
            self.kernel = kernel
        else:                                                            # kriging.py:85-88
            l_scale = (1.0,) * dim
            self.scale_bounds = [(0.01, 100)] * dim
            self.kernel = ConstantKernel() * Matern(length_scale=l_scale, length_scale_bounds=self.scale_bounds)
        if noise:                                                        # kriging.py:91-96
            if isinstance(noise, bool):
                noise = WhiteKernel()
            else:
                noise = WhiteKernel(noise_level=noise)
            self.kernel += noise
        if global_optimizer:                                             # kriging.py:100-106
            self.n_restart = 3 if n_restart is None else int(n_restart)
            if self.n_restart < 1:
                raise ValueError("n_restart must be >= 1")
            n_local = 0
        else:
            self.n_restart = 1
            n_local = 10 * dim
        self.n_cpu = 1    # the fork pools of kriging.py:109-115 are replaced by batched device calls

        # fail before any fit work if the kernel tree has no device implementation
        kernel_spec.flatten(self.kernel, dim)

        import torch
        from . import fit as _fit
        if not torch.cuda.is_available():
            raise RuntimeError("batman_b200.Kriging needs a CUDA device: there is no CPU fallback")
        X_dev = torch.from_numpy(np.ascontiguousarray(sample)).cuda()
        gps, hyper, wts = [], [], []
        fitted, self.fit_stats = _fit.fit_outputs(self.kernel, X_dev, sample, data, global_optimizer, n_local,
                                                  self.n_restart)
        for kernel_, _, alpha, L, y_mean, y_std, lml, wt in fitted:
            hyperparameter = np.exp(kernel_.theta)                       # kriging.py:123
            if kernel is None:                                           # kriging.py:126-134
                hyper_bounds = all([i[0] < j < i[1] for i, j in zip(self.scale_bounds, hyperparameter[1:dim + 1])])
                if not hyper_bounds:
                    self.logger.warning("Hyperparameters optimization not converged: {}".format(kernel_))
            gps.append(FittedGP(kernel_, sample, alpha, L, y_mean, y_std, lml))
            hyper.append(hyperparameter)
            wts.append(wt)
        self.gp_models, self.hyperparameter = tuple(gps), tuple(hyper)   # kriging.py:148
        self.logger.debug("Kernels:\n{}".format([gp.kernel_ for gp in self.gp_models]))
        self._device = None
        self._wts = wts            # packed L^-1 from the fit (device tensors; dropped on pickling)

    # ------------------------------------------------------------------ alternate constructors
    @classmethod
    def from_state(cls, gp_models, kernel=None):
        """Wrap already-fitted per-output states (objects with ``kernel_, X_train_, alpha_, L_,
        _y_train_mean, _y_train_std`` -- e.g. fitted ``GaussianProcessRegressor``s, which is what
        ``batman.surrogate.Kriging.gp_models`` holds).  No arithmetic: this is how parity on
        *identical* fitted state is tested, and how a reference pickle is adopted."""
        self = cls.__new__(cls)
        gps = []
        for gp in gp_models:
            ym, ys = np.ravel(gp._y_train_mean), np.ravel(gp._y_train_std)
            gps.append(FittedGP(gp.kernel_, np.asarray(gp.X_train_, dtype=np.float64),
                                np.asarray(gp.alpha_, dtype=np.float64).ravel(),
                                np.asarray(gp.L_, dtype=np.float64), float(ym[0]), float(ys[0]),
                                getattr(gp, 'log_marginal_likelihood_value_', None)))
        self.gp_models = tuple(gps)
        self.model_len = len(gps)
        self.kernel = kernel if kernel is not None else gps[0].kernel_
        self.hyperparameter = tuple(np.exp(gp.kernel_.theta) for gp in gps)
        self.n_restart, self.n_cpu = 1, 1
        self._device = None
        self._wts = None
        return self

    # ------------------------------------------------------------------ device state (lazy, per process)
    def _model(self):
        if self._device is None:
            from .device_model import DeviceModel
            X = self.gp_models[0].X_train_
            dim = X.shape[1]
            specs = [kernel_spec.flatten(gp.kernel_, dim) for gp in self.gp_models]
            self._device = DeviceModel(X, specs, [gp.alpha_ for gp in self.gp_models],
                                       [gp.L_ for gp in self.gp_models],
                                       [gp._y_train_mean for gp in self.gp_models],
                                       [gp._y_train_std for gp in self.gp_models],
                                       wts=getattr(self, '_wts', None))
        return self._device

    @contextlib.contextmanager
    def input_scaler(self, scale, mn):
        """Fuse ``x*scale + min`` (MinMaxScaler.transform, surrogate_model.py:181) into the kernels for the
        duration of ONE call, so that the facade can feed raw points; on exit the predictor is back to
        ``evaluate``'s own contract (already-scaled input, kriging.py:192) -- the scaler is never object state."""
        model = self._model()
        model.set_scaler(np.asarray(scale, float), np.asarray(mn, float))
        try:
            yield model
        finally:
            model.set_scaler(None, None)

    def __getstate__(self):
        """dill/pickle (surrogate_model.py:271-279): numpy state only, device buffers are rebuilt lazily."""
        state = self.__dict__.copy()
        state['_device'] = None
        state['_wts'] = None
        state.pop('_scaler', None)          # pickles of the first release carried a sticky scaler
        return state

    # ------------------------------------------------------------------ evaluate
    def evaluate(self, point, return_std=True):
        """Make a prediction (kriging.py:192-214).

        :param array_like point: point(s) to evaluate, (n_features,) or
          (n_points, n_features).
        :return: predictions and standard deviations, each
          (n_points, n_outputs).
        """
        x_n = np.atleast_2d(np.asarray(point, dtype=np.float64))         # functions/utils.py:19
        n_features = self.gp_models[0].X_train_.shape[1]
        if x_n.ndim != 2 or x_n.shape[1] != n_features:
            raise ValueError("X has %d features, but Kriging is expecting %d features as input."
                             % (x_n.shape[-1], n_features))              # sklearn's validate_data message
        mean, sigma = self._model().predict(x_n, return_std=return_std)
        if not return_std:
            return mean
        return mean, sigma
