"""GPU tests of the callers around the hot path: Evofusion, leave-one-out quality,
sigma / expected-improvement resampling criteria.  Parity is stated at FIXED hyper-parameters against scikit-learn
(the reference's own fits are unseeded DE runs, kriging.py:170-171); the reference's functional tests
(tests/test_models.py:253-268 Evofusion on Forrester's functions) are mirrored with their tolerances."""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process import kernels as sk
from sklearn.metrics import r2_score

from oracle import functions as ofun

pytestmark = pytest.mark.gpu


def _forrester(x, fidelity):
    """functions/analytical.py:392-405."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    f_e = (6 * x - 2) ** 2 * np.sin(12 * x - 4)
    return f_e if fidelity == 'e' else 0.5 * f_e + 10 * (x - 0.5) - 5


def _mufi(n_e=5, n_c=13):
    x_c = ofun.halton(n_c, 1)
    x_e = x_c[:n_e]
    sample = np.vstack([np.hstack([np.zeros((n_e, 1)), x_e]), np.hstack([np.ones((n_c, 1)), x_c])])
    data = np.vstack([_forrester(x_e, 'e'), _forrester(x_c, 'c')])
    return sample, data, x_e, x_c


def test_evofusion_fixed_theta_equals_sum_of_sklearn_models():
    from batman_b200 import Evofusion
    sample, data, x_e, x_c = _mufi()
    kern = sk.ConstantKernel(1.0, 'fixed') * sk.Matern([0.3], 'fixed', nu=2.5)
    evo = Evofusion(sample, data, kernel=kern)
    gc = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(x_c, _forrester(x_c, 'c')[:, 0])
    ge = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(
        x_e, (_forrester(x_e, 'e') - _forrester(x_e, 'c'))[:, 0])
    pts = np.linspace(0, 1, 201).reshape(-1, 1)
    pred, sigma = evo.evaluate(pts)
    mc, sc = gc.predict(pts, return_std=True)
    me, se = ge.predict(pts, return_std=True)
    np.testing.assert_allclose(pred[:, 0], mc + me, rtol=0, atol=1e-8 * np.abs(mc + me).max())
    np.testing.assert_allclose(sigma[:, 0] ** 2, (sc + se) ** 2, rtol=0, atol=1e-8 * (sc + se).max() ** 2 + 1e-12)


def test_evofusion_forrester_like_the_reference_test():
    """tests/test_models.py:253-268: point prediction within 10 %, Q2 within 10 % of 1."""
    from batman_b200 import Evofusion
    np.random.seed(123456)
    sample, data, _, _ = _mufi(n_e=10, n_c=15)      # Space(..., multifidelity=[5.1, 13.0]), doe_init 10
    evo = Evofusion(sample, data)
    pred, _ = evo.evaluate([0.65])
    assert pred.shape == (1, 1)
    assert pred[0, 0] == pytest.approx(_forrester([0.65], 'e')[0, 0], rel=0.1)
    x = np.random.RandomState(0).rand(1000, 1)
    q2 = r2_score(_forrester(x, 'e'), evo.evaluate(x)[0])
    assert q2 == pytest.approx(1, abs=0.1)


def test_estimate_quality_loo_matches_sklearn_at_fixed_theta():
    """surrogate_model.py:200-269: n refits, r2_score of the held-out predictions, point of largest error."""
    from batman_b200 import SurrogateModel
    n, p = 24, 2
    corners = np.array([[-1.0, 0.0], [3.0, 5.0]])
    X = corners[0] + ofun.halton(n, p) * (corners[1] - corners[0])
    y = (np.sin(X[:, :1]) * X[:, 1:] + 0.3 * X[:, :1] ** 2)
    kern = sk.ConstantKernel(2.0, 'fixed') * sk.Matern([0.4, 0.6], 'fixed', nu=1.5)
    s = SurrogateModel('kriging', corners, None, kernel=kern)
    s.fit(X, y)
    q2, point = s.estimate_quality()
    Xs = (X - corners[0]) / (corners[1] - corners[0])
    pred = np.empty(n)
    for i in range(n):
        keep = np.arange(n) != i
        pred[i] = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(
            Xs[keep], y[keep, 0]).predict(Xs[i:i + 1])[0]
    assert q2 == pytest.approx(r2_score(y[:, 0], pred), abs=1e-8)
    assert np.allclose(point, X[np.argmax((y[:, 0] - pred) ** 2)])


def test_refiner_sigma_and_expected_improvement():
    """refiner.py:287-325, 523-592 with the DE population batched through the device."""
    from batman_b200 import Refiner, SurrogateModel
    np.random.seed(123456)
    n, p = 10, 2       # few points: the reference's closeness test excludes bands of +-0.02 around every coordinate
    corners = np.array([[0.0, 0.0], [1.0, 1.0]])
    X = ofun.halton(n, p)
    y = ofun.g_function(X)[:, :1]
    kern = sk.ConstantKernel(1.0, 'fixed') * sk.Matern([0.3, 0.3], 'fixed', nu=1.5)
    s = SurrogateModel('kriging', corners, None, kernel=kern)
    s.fit(X, y)
    r = Refiner(s, corners)
    dense = r._lo + ofun.halton(20000, p) * (r._hi - r._lo)
    _, sig = s(dense)
    x_sigma = r.sigma()
    assert np.all(x_sigma >= r._lo) and np.all(x_sigma <= r._hi)
    assert s(x_sigma)[1][0, 0] >= 0.98 * sig.max()
    # expected improvement: at least as good as the best of the dense sample (away from the existing points)
    x_ei = r.optimization(method='EI', extremum='min')
    mean, sigma = s(dense)
    min_value = s(X)[0].min()
    from scipy.stats import norm
    diff = min_value - mean[:, 0]
    std = np.sqrt(sigma[:, 0])
    ei = diff * norm.cdf(diff / std) + std * norm.pdf(diff / std)
    xs = (dense - r._lo) / (r._hi - r._lo)
    far = ~(np.abs(xs[:, None, :] - r.points_scaled[None, :, :]).min(axis=2) < 0.02).any(axis=1)
    assert far.sum() > 1000
    m1, s1 = s(x_ei)
    d1 = min_value - m1[0, 0]
    ei1 = d1 * norm.cdf(d1 / np.sqrt(s1[0, 0])) + np.sqrt(s1[0, 0]) * norm.pdf(d1 / np.sqrt(s1[0, 0]))
    assert ei1 >= 0.9 * ei[far].max()


def test_mixture_fixed_theta_equals_manual_pipeline():
    """surrogate/mixture.py:100-205, 305-353: cluster -> classify -> local Kriging, against the same pipeline
    assembled by hand from scikit-learn objects at a fixed kernel."""
    from sklearn.cluster import KMeans
    from sklearn.svm import SVC
    from batman_b200 import Mixture
    rng = np.random.RandomState(5)
    corners = np.array([[0.0, 0.0], [2.0, 1.0]])
    X = corners[0] + ofun.halton(60, 2) * (corners[1] - corners[0])
    step = (X[:, :1] > 1.0).astype(float)
    data = np.hstack([np.sin(3 * X[:, :1]) + 10 * step + X[:, 1:], np.cos(2 * X[:, 1:]) - 10 * step])
    kern = sk.ConstantKernel(1.0, 'fixed') * sk.Matern([0.4, 0.6], 'fixed', nu=2.5)
    local = [{'kriging': {'kernel': kern}}] * 2
    mix = Mixture(X, data, corners, local_method=local, clusterer=KMeans(n_clusters=2, n_init=10, random_state=0),
                  classifier=SVC())
    assert sorted(len(v) for v in mix.indice_clt.values()) == sorted([int(step.sum()), int(60 - step.sum())])
    pts = corners[0] + rng.rand(300, 2) * (corners[1] - corners[0])
    pred, sigma, classif = mix.evaluate(pts, classification=True)
    assert pred.shape == (300, 2) and sigma.shape == (300, 2)
    Xs = (X - corners[0]) / (corners[1] - corners[0])
    ps = (pts - corners[0]) / (corners[1] - corners[0])
    want_m, want_s = np.empty_like(pred), np.empty_like(sigma)
    for k in mix.clusters_id:
        idx = mix.indice_clt[k]
        sel = classif == k
        for q in range(2):
            gp = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(Xs[idx], data[idx, q])
            want_m[sel, q], want_s[sel, q] = gp.predict(ps[sel], return_std=True)
    np.testing.assert_allclose(pred, want_m, rtol=0, atol=1e-8 * np.abs(want_m).max())
    np.testing.assert_allclose(sigma ** 2, want_s ** 2, rtol=0, atol=1e-8 * (want_s ** 2).max() + 1e-12)
    assert (classif == mix.classifier.predict(ps)).all()


def test_mixture_docstring_example_runs():
    """surrogate/mixture.py:11-29 (default clusterer / classifier strings, default local Kriging with DE fits).  The
    docstring's numbers come from an unseeded KMeans + unseeded DE, so only the contract is asserted."""
    from batman_b200 import Mixture
    np.random.seed(123456)
    samples = np.array([[1., 5.], [2., 5.], [8., 5.], [9., 5.]])
    data = np.array([[50., 51., 52.], [49., 48., 47.], [10., 11., 12.], [9., 8., 7.]])
    sample_new = np.array([[0.5, 5.], [10., 5.], [8.5, 5.]])
    corners = np.array([[1., 5.], [9., 5.]])
    algo = Mixture(samples, data, corners, 3)
    pred, sigma = algo.evaluate(sample_new)
    print(pred, sigma)
    assert pred.shape == (3, 3) and sigma.shape == (3, 3)
    assert np.all(np.isfinite(pred)) and np.all(sigma >= 0)
    with pytest.raises(AttributeError):
        Mixture(samples, data, corners, 3, clusterer='cluster.NoSuchThing()')
