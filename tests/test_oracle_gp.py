"""CPU: pin the oracle restatement against the reference's golden vectors and sklearn."""
import numpy as np
import pytest
from scipy.spatial.distance import cdist

from conftest import GP_CASES, load_golden
from oracle import gp as ogp
from oracle import gp_hp


def test_docstring_known_answer():
    """kriging.py:14-19: evaluate((5, 8)) -> (10.333, 3.591), (1.247, 0.694)."""
    z, states = load_golden('kriging_docstring')
    mean, sigma = ogp.evaluate(states, z['points'])
    np.testing.assert_allclose(mean[0], [10.333, 3.591], atol=5e-4)
    np.testing.assert_allclose(sigma[0], [1.247, 0.694], atol=5e-4)
    np.testing.assert_allclose(z['ref_mean'][0], [10.333, 3.591], atol=5e-4)


@pytest.mark.parametrize('case', GP_CASES)
def test_oracle_matches_reference_evaluate(case):
    """oracle.gp.evaluate == batman.surrogate.Kriging.evaluate on the dumped state."""
    z, states = load_golden(case)
    mean, sigma = ogp.evaluate(states, z['points'])
    scale = max(s['y_std'] for s in states)
    assert mean.shape == z['ref_mean'].shape
    np.testing.assert_allclose(mean, z['ref_mean'], rtol=0, atol=1e-11 * max(1.0, np.abs(z['ref_mean']).max()))
    np.testing.assert_allclose(sigma, z['ref_sigma'], rtol=0, atol=1e-8 * scale)
    # variance is the well-conditioned quantity (sigma = sqrt amplifies near 0)
    np.testing.assert_allclose(sigma ** 2, z['ref_sigma'] ** 2, rtol=0, atol=1e-12 * scale ** 2)


def test_ishigami_through_scaler():
    """SurrogateModel.__call__ = MinMax scaling (surrogate_model.py:181) + Kriging.evaluate."""
    z, states = load_golden('ishigami_n100')
    lo, hi = z['corners']
    pts = (z['points'] - lo) / (hi - lo)
    mean, sigma = ogp.evaluate(states, pts)
    np.testing.assert_allclose(mean, z['ref_mean'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(sigma ** 2, z['ref_sigma'] ** 2, rtol=0, atol=1e-11 * states[0]['y_std'] ** 2)


def test_cdist_is_sequential_non_fused():
    rng = np.random.RandomState(0)
    for p in (1, 2, 3, 8, 20):
        U, V = rng.rand(64, p) * 3, rng.rand(80, p) * 3
        assert np.array_equal(cdist(U, V), ogp._scaled_dists(U, V, np.ones(p)))


def test_fit_state_matches_sklearn_factorisation():
    z, states = load_golden('ishigami_n100')
    st = states[0]
    mine = ogp.fit_state(st['tree'], st['X_train'], z['y'].ravel())
    np.testing.assert_allclose(mine['L'], st['L'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(mine['alpha'], st['alpha'], rtol=1e-7)
    assert abs(mine['y_mean'] - st['y_mean']) < 1e-14 and abs(mine['y_std'] - st['y_std']) < 1e-14


def test_lml_matches_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    rng = np.random.RandomState(1)
    X = rng.rand(60, 3)
    y = np.sin(4 * X[:, 0]) + X[:, 1] * X[:, 2]
    kern = ConstantKernel(2.0) * Matern(length_scale=[0.5, 1.0, 2.0], nu=1.5)
    gp = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, y)
    yn, _, _ = ogp.normalise_y(y)
    for theta in [kern.theta, kern.theta + 0.3, np.log([1e-3, 50.0, 50.0, 50.0])]:
        ref = gp.log_marginal_likelihood(theta)
        mine = ogp.lml(ogp.from_sklearn(kern.clone_with_theta(theta)), X, yn)
        assert mine == pytest.approx(ref, rel=1e-12, abs=1e-9)


@pytest.mark.parametrize('case', ['kernel_default_3out', 'kernel_const_times_rbf_ard'])
def test_longdouble_arbiter_close_to_float64(case):
    z, states = load_golden(case)
    for st in states:
        m64, s64 = ogp.predict(st, z['points'])
        mld, sld = gp_hp.predict_hp(st, z['points'])
        assert np.abs(m64 - mld.astype(float)).max() < 1e-10 * max(1.0, st['y_std'])
        assert np.abs(s64 ** 2 - (sld ** 2).astype(float)).max() < 1e-11 * st['y_std'] ** 2
