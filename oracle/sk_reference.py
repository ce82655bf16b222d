"""The reference call sequence of ``batman.surrogate.Kriging`` on scikit-learn.

TEST INFRASTRUCTURE ONLY.  This is the "reference CPU path" that bench.py
times (``--impl reference`` and ``cpu_baseline``): batman itself cannot travel
to the GPU box (`/root/reference` does not exist there), but the library that
does all of its arithmetic -- scikit-learn -- is installed in the image, so the
path is restated here call for call:

* construction   kriging.py:85-106  (default kernel, WhiteKernel noise,
                 normalize_y=True, optimizer choice)
* fit            kriging.py:117-123 (one GaussianProcessRegressor per column)
* evaluate       kriging.py:192-214 wrapped by functions/utils.py:19-31
                 (python loop, one point per ``gp.predict`` call)
* batched        the strongest fair CPU baseline: the same
                 fitted regressors called on blocks of points, BLAS on all cores.
"""
import warnings

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel


def default_kernel(dim):
    """kriging.py:85-88."""
    return ConstantKernel() * Matern(length_scale=(1.0,) * dim,
                                     length_scale_bounds=[(0.01, 100)] * dim)


class SkKriging:
    """Reference Kriging without the fork pools (those only parallelise fit)."""

    def __init__(self, sample, data, kernel=None, noise=False, optimizer='fmin_l_bfgs_b',
                 n_restarts_optimizer=0, random_state=None):
        sample = np.atleast_2d(sample)
        data = np.asarray(data, dtype=np.float64)
        dim = sample.shape[1]
        self.model_len = data.shape[1]                       # kriging.py:78
        self.kernel = kernel if kernel is not None else default_kernel(dim)
        if noise:                                            # kriging.py:91-96
            self.kernel = self.kernel + (WhiteKernel() if isinstance(noise, bool)
                                         else WhiteKernel(noise_level=noise))
        self.gp_models = []
        for q in range(self.model_len):
            gp = GaussianProcessRegressor(kernel=self.kernel, normalize_y=True, optimizer=optimizer,
                                          n_restarts_optimizer=n_restarts_optimizer,
                                          random_state=random_state)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                gp.fit(sample, data[:, q])                   # kriging.py:121 (column is 1-D)
            self.gp_models.append(gp)
        self.hyperparameter = [np.exp(gp.kernel_.theta) for gp in self.gp_models]

    def evaluate_point(self, point):
        """Body of Kriging.evaluate, kriging.py:201-214."""
        point_array = np.atleast_2d(point)
        prediction = np.empty((self.model_len))
        sigma = np.empty((self.model_len))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i, gp in enumerate(self.gp_models):
                pred, sig = gp.predict(point_array, return_std=True, return_cov=False)
                prediction[i], sigma[i] = pred[0], sig[0]
        return prediction, sigma

    def evaluate(self, points):
        """multi_eval wrapper, functions/utils.py:19-31: one python iteration per point."""
        x_n = np.atleast_2d(points)
        feval = [self.evaluate_point(x_i) for x_i in x_n]
        feval, sigma = zip(*feval)
        return (np.array(feval).reshape((len(x_n), -1)),
                np.array(sigma).reshape((len(x_n), -1)))

    def evaluate_batched(self, points, chunk=8192, return_std=True):
        """Same fitted models, block calls (not what batman does; the fair CPU bar)."""
        x_n = np.atleast_2d(points)
        mean = np.empty((len(x_n), self.model_len))
        sigma = np.empty_like(mean) if return_std else None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for s in range(0, len(x_n), chunk):
                for i, gp in enumerate(self.gp_models):
                    if return_std:
                        mean[s:s + chunk, i], sigma[s:s + chunk, i] = gp.predict(
                            x_n[s:s + chunk], return_std=True)
                    else:
                        mean[s:s + chunk, i] = gp.predict(x_n[s:s + chunk])
        return (mean, sigma) if return_std else mean


def fixed_theta_kriging(sample, data, kernels):
    """Per-output regressors with *given* fitted kernels (optimizer=None): the
    final factorisation of sklearn:_gpr.py:349-367 only.  Used to build large
    deterministic models (n_train 1024..16384) without a multi-hour DE fit."""
    obj = SkKriging.__new__(SkKriging)
    sample = np.atleast_2d(sample)
    data = np.asarray(data, dtype=np.float64)
    obj.model_len = data.shape[1]
    obj.kernel = kernels[0]
    obj.gp_models = []
    for q in range(obj.model_len):
        gp = GaussianProcessRegressor(kernel=kernels[q], normalize_y=True, optimizer=None)
        gp.fit(sample, data[:, q])
        obj.gp_models.append(gp)
    obj.hyperparameter = [np.exp(gp.kernel_.theta) for gp in obj.gp_models]
    return obj
