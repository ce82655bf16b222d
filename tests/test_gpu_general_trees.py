"""GPU parity for general kernel trees (several stationary leaves; SURVEY.md §8 A5 note).

The reference accepts any ``sklearn.gaussian_process.kernels`` expression (kriging.py:80-96,
driver.py:181-190).  Trees with one stationary leaf run on the fast kernels (test_gpu_predict.py);
Sum/Product trees over up to three RBF/Matern leaves run on the general kernels
(csrc/gp_predict_gen.cu, k_build_K, the unfused Saltelli composition).  Same gates as everywhere:
mean/variance 1e-10 relative to the output scale, Sobol' indices 1e-8 absolute, LML vs scikit-learn.
"""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process import kernels as sk

from conftest import kriging_from_states, tree_to_sklearn
from oracle import functions as ofun
from oracle import gp as ogp
from oracle import philox, saltelli

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _trees(p):
    ls_a = np.linspace(0.6, 1.4, p)
    ls_b = np.linspace(1.5, 0.5, p)
    ls_c = np.full(p, 2.0)
    return {
        'c*rbf*matern52': ('prod', ('prod', ('const', 2.0), ('rbf', ls_a)), ('matern', 2.5, ls_b)),
        'c1*rbf+c2*matern12+white': ('sum', ('sum', ('prod', ('const', 1.5), ('rbf', ls_a)),
                                             ('prod', ('const', 0.5), ('matern', 0.5, ls_b))), ('white', 1e-3)),
        '(c+rbf)*(c+matern32)': ('prod', ('sum', ('const', 0.3), ('rbf', ls_a)), ('sum', ('const', 0.2), ('matern', 1.5, ls_b))),
        'rbf*rbf': ('prod', ('rbf', ls_a), ('rbf', ls_a)),
        'three leaves': ('sum', ('sum', ('matern', 0.5, ls_a), ('prod', ('const', 0.7), ('matern', 1.5, ls_b))),
                         ('matern', np.inf, ls_c)),
    }


def _state(tree, n, p, noise=0.0):
    X = ofun.halton(n, p)
    y = ofun.g_function(X)[:, 0] + noise * np.random.RandomState(0).randn(n)
    return ogp.fit_state(tree, X, y), X, y


@pytest.mark.parametrize('name', list(_trees(4)))
@pytest.mark.parametrize('n,p', [(300, 4), (1000, 16)])
def test_general_tree_predict_parity(name, n, p):
    from batman_b200 import kernel_spec
    tree = _trees(p)[name]
    spec = kernel_spec.flatten(tree_to_sklearn(tree), p)
    assert spec.general is not None
    st, X, y = _state(tree, n, p)
    k = kriging_from_states([st])
    rng = np.random.RandomState(7)
    pts = np.vstack([rng.rand(700, p), X[:3], X[:3] + 1e-7])
    omean, osigma = ogp.evaluate([st], pts)
    prior = ogp.kernel_diag(tree) * st['y_std'] ** 2
    scale = np.maximum(np.abs(omean[:, 0]), st['y_std'])
    # independent pin: scikit-learn itself on the same data and the same fixed theta
    ref = GaussianProcessRegressor(kernel=tree_to_sklearn(tree), normalize_y=True, optimizer=None).fit(X, y)
    rm, rs = ref.predict(pts, return_std=True)
    amp = np.abs(ref.kernel_(pts, X)) @ np.abs(ref.alpha_) * st['y_std'] / scale        # conditioning of the mean sum
    gate = np.maximum(RTOL, 8 * np.finfo(float).eps * amp)
    for kernel in ('int8', 'dmma'):
        k._model().set_sigma_kernel(kernel)
        mean, sigma = k.evaluate(pts)
        assert np.all(np.abs(mean[:, 0] - omean[:, 0]) / scale <= gate)
        assert np.all(np.abs(mean[:, 0] - rm) / scale <= gate)
        assert np.max(np.abs(sigma[:, 0] ** 2 - osigma[:, 0] ** 2)) <= RTOL * prior
        assert np.max(np.abs(sigma[:, 0] ** 2 - rs ** 2)) <= 1e-9 * prior
    mean_only = k._model().predict(pts, return_std=False)[0]
    assert np.array_equal(mean_only, mean)


def test_general_tree_lml_matches_sklearn():
    import torch
    from batman_b200 import kernel_spec
    from batman_b200.fit import normalise_y
    from batman_b200.lml import lml_batched
    n, p = 200, 3
    X = ofun.halton(n, p)
    y = ofun.g_function(X)[:, 0] + 0.01 * np.random.RandomState(0).randn(n)
    yn, _, _ = normalise_y(y)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda()
    rng = np.random.RandomState(2)
    for kern in [sk.ConstantKernel(1.0) * sk.RBF(length_scale=[1.0] * p) * sk.Matern(length_scale=[0.8] * p, nu=2.5),
                 sk.ConstantKernel(1.5) * sk.RBF(length_scale=1.0) + sk.ConstantKernel(0.5) * sk.Matern(length_scale=2.0, nu=0.5)
                 + sk.WhiteKernel(0.01)]:
        gp = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, y)
        thetas = [kern.theta] + [kern.theta + rng.uniform(-1.0, 1.0, size=len(kern.theta)) for _ in range(5)]
        want = np.array([gp.log_marginal_likelihood(t) for t in thetas])
        got = lml_batched(Xd, yd, [kernel_spec.flatten(kern.clone_with_theta(t), p) for t in thetas])
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-7)
        mine = np.array([ogp.lml(ogp.from_sklearn(kern.clone_with_theta(t)), X, yn) for t in thetas])
        np.testing.assert_allclose(got, mine, rtol=1e-9, atol=1e-7)


def test_general_tree_kriging_fit_state_and_optimiser():
    """Kriging(sample, data, kernel=<general tree>): final factorisation vs scikit-learn at a fixed theta, and the
    L-BFGS-B path (kriging.py:99-106 with global_optimizer=False) reaches at least scikit-learn's optimum."""
    from batman_b200 import Kriging
    n, p = 150, 3
    X = ofun.halton(n, p)
    y = ofun.g_function(X)[:, 0]
    fixed = sk.ConstantKernel(1.7, 'fixed') * sk.RBF([1.0, 1.5, 2.0], 'fixed') \
        + sk.ConstantKernel(0.4, 'fixed') * sk.Matern([0.5, 0.6, 0.7], 'fixed', nu=1.5)
    k = Kriging(X, y[:, None], kernel=fixed)
    ref = GaussianProcessRegressor(kernel=fixed, normalize_y=True, optimizer=None).fit(X, y)
    gp = k.gp_models[0]
    np.testing.assert_allclose(gp.L_, ref.L_, rtol=0, atol=1e-11)
    assert gp.log_marginal_likelihood_value_ == pytest.approx(ref.log_marginal_likelihood_value_, rel=1e-10, abs=1e-8)
    pts = np.random.RandomState(3).rand(400, p)
    mean, sigma = k.evaluate(pts)
    rm, rs = ref.predict(pts, return_std=True)
    np.testing.assert_allclose(mean[:, 0], rm, rtol=0, atol=1e-8 * np.abs(rm).max())
    np.testing.assert_allclose(sigma[:, 0] ** 2, rs ** 2, rtol=0, atol=1e-9 * ref._y_train_std ** 2 * 2.1)
    # free hyper-parameters, local optimiser
    np.random.seed(123456)
    free = sk.ConstantKernel(1.0) * sk.RBF([1.0] * p, (0.05, 20.0)) + sk.ConstantKernel(0.1) * sk.Matern([1.0] * p, (0.05, 20.0), nu=0.5)
    kf = Kriging(X, y[:, None], kernel=free, global_optimizer=False)
    skl = GaussianProcessRegressor(kernel=free, normalize_y=True, n_restarts_optimizer=0).fit(X, y)
    assert kf.gp_models[0].log_marginal_likelihood_value_ >= skl.log_marginal_likelihood_value_ - 1e-3 * abs(
        skl.log_marginal_likelihood_value_)
    mean, _ = kf.evaluate(X)
    assert np.max(np.abs(mean[:, 0] - y)) < 1e-3 * np.ptp(y)


def test_general_tree_sobol_vs_oracle_same_design():
    """UQ.sobol() with a general-tree output next to a fast-path output (m = 2): the unfused composition fills the
    same partial slots; estimator parity 1e-8 on identical base samples; in-kernel Philox == explicit samples."""
    from batman_b200 import SurrogateModel, UQ
    n, p, N = 200, 3, 1500
    X = ofun.halton(n, p)
    Y = np.column_stack([ofun.g_function(X)[:, 0], ofun.ishigami(2 * np.pi * X - np.pi)[:, 0]])
    tree_g = _trees(p)['c1*rbf+c2*matern12+white']
    tree_a = ('prod', ('const', 1.5), ('matern', 1.5, np.full(p, 0.8)))
    states = [ogp.fit_state(tree_g, X, Y[:, 0]), ogp.fit_state(tree_a, X, Y[:, 1])]
    corners = np.array([[0.0] * p, [1.0] * p])
    s = SurrogateModel('kriging', corners, None)
    s.adopt(kriging_from_states(states), X)
    dists = [('uniform', 0.0, 1.0)] * p
    A, B = philox.base_samples(123456, 0, N, dists)
    uq = UQ(s, dists=['Uniform(0., 1.)'] * p, nsample=N, base_samples=(A, B))
    s2, s1, st = uq.sobol()
    Yd = ogp.evaluate(states, saltelli.design_from_base(A, B))[0]
    want = saltelli.uq_sobol_aggregate(Yd, N, p)
    np.testing.assert_allclose(s1, want[1], rtol=0, atol=1e-8)
    np.testing.assert_allclose(st, want[2], rtol=0, atol=1e-8)
    np.testing.assert_allclose(s2, want[0], rtol=0, atol=1e-8)
    est = saltelli.saltelli_indices(Yd, N, p)
    np.testing.assert_allclose(uq.details['S1'], est['S1'], rtol=0, atol=1e-8)
    np.testing.assert_allclose(uq.details['ST'], est['ST'], rtol=0, atol=1e-8)
    g2, g1, gt = UQ(s, dists=['Uniform(0., 1.)'] * p, nsample=N, seed=123456).sobol()
    np.testing.assert_allclose(g1, s1, rtol=0, atol=1e-13)
    np.testing.assert_allclose(gt, st, rtol=0, atol=1e-13)
    np.testing.assert_allclose(g2, s2, rtol=0, atol=1e-13)


def test_general_tree_block_indices_store_path():
    """indices='block' stores the per-block outputs (bk_sobol_modes) before the trapz reduction: the general
    composition must fill the same [block][row][output] layout, next to a fast-path output; in-kernel Philox rows
    beyond one slab (2^18 base rows) exercise the slab loop."""
    from batman_b200 import SurrogateModel, UQ
    from oracle import saltelli as osal
    n, p = 120, 3
    X = ofun.halton(n, p)
    Y = np.column_stack([ofun.g_function(X)[:, 0], ofun.ishigami(2 * np.pi * X - np.pi)[:, 0],
                         np.sin(3 * X[:, 0]) + X[:, 1] * X[:, 2]])
    trees = [_trees(p)['c*rbf*matern52'], ('prod', ('const', 1.5), ('matern', 1.5, np.full(p, 0.8))),
             _trees(p)['three leaves']]
    states = [ogp.fit_state(t, X, Y[:, q]) for q, t in enumerate(trees)]
    s = SurrogateModel('kriging', np.array([[0.0] * p, [1.0] * p]), None)
    s.adopt(kriging_from_states(states), X)
    N = 1000
    A, B = philox.base_samples(11, 0, N, [('uniform', 0.0, 1.0)] * p)
    got = UQ(s, dists=['Uniform(0., 1.)'] * p, nsample=N, indices='block', base_samples=(A, B)).sobol()
    Yd = osal.int_func(ogp.evaluate(states, osal.design_from_base(A, B))[0])
    want = osal.uq_sobol_aggregate(Yd, N, p, indices='block')
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=0, atol=1e-8)
    # slab loop of the general composition: 300 000 base rows > 2^18, generated in-kernel, vs explicit samples
    Nbig = 300000
    one = SurrogateModel('kriging', np.array([[0.0] * p, [1.0] * p]), None)
    small = ogp.fit_state(_trees(p)['c1*rbf+c2*matern12+white'], X[:32], Y[:32, 0])
    one.adopt(kriging_from_states([small]), X[:32])
    Ab, Bb = philox.base_samples(5, 0, Nbig, [('uniform', 0.0, 1.0)] * p)
    r_gen = UQ(one, dists=['Uniform(0., 1.)'] * p, nsample=Nbig, seed=5).sobol()
    r_exp = UQ(one, dists=['Uniform(0., 1.)'] * p, nsample=Nbig, base_samples=(Ab, Bb)).sobol()
    for g, w in zip(r_gen, r_exp):
        np.testing.assert_allclose(g, w, rtol=0, atol=1e-12)


def test_general_tree_limits_raise():
    from batman_b200 import Kriging
    p = 20                                                  # 3 leaves x padded dimension 20 > 48
    X = ofun.halton(64, p)
    y = ofun.g_function(X)[:, :1]
    kern = sk.Matern([1.0] * p, 'fixed', nu=0.5) + sk.RBF([2.0] * p, 'fixed') + sk.Matern([3.0] * p, 'fixed', nu=2.5)
    with pytest.raises(NotImplementedError):
        Kriging(X, y, kernel=kern)
