"""GPU parity at the sizes BASELINE.json's configs are quoted on (the ones round 1 only timed).

* cfg2  G-function 8-D, n_train = 1024: UQ.sobol() with N = 1e5 base rows (1.8e6 predictions) generated
        IN-KERNEL, against the oracle on the same Philox rows -- Sobol' indices within 1e-8.
* cfg4  Channel_Flow field F = 400, 20 POD modes (128 Halton snapshots), near-singular trailing modes:
        SurrogateModel.__call__ (mean and sigma through the POD) and UQ.sobol() (aggregated + map) vs the oracle.
* cfg5  synthetic 20-D anisotropic RBF, n_train = 16384 (n_pad > 8192: the 8 x 7-bit int8 slicing is the only exact
        one): mean + sigma on BOTH sigma kernels vs the oracle and the long-double arbiter.
* cfg3  log-marginal likelihood at n_train = 4096 (Michalewicz 10-D) vs scikit-learn.

Tolerances are the north_star's: mean / variance 1e-10 of the output scale / prior variance, indices 1e-8.
"""
import numpy as np
import pytest

from conftest import kriging_from_states, synthetic_state
from oracle import functions as ofun
from oracle import gp as ogp
from oracle import gp_hp, philox, pod as opod, saltelli

pytestmark = pytest.mark.gpu

RTOL = 1e-10
ATOL_S = 1e-8
EPS = np.finfo(np.float64).eps


def _oracle_means(states, X, chunk=16384):
    """oracle/gp.py predict (mean only) over a large design: bounded slabs, one host thread per slab (cdist / exp
    release the GIL), so that 1.8e6 oracle predictions take tens of seconds on the box's cores."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    out = np.empty((len(X), len(states)))

    def slab(lo):
        for q, st in enumerate(states):
            out[lo:lo + chunk, q] = ogp.predict(st, X[lo:lo + chunk], return_std=False)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(slab, range(0, len(X), chunk)))
    return out


# --------------------------------------------------------------------------------------------- cfg2
def test_cfg2_sobol_N1e5_in_kernel_design_vs_oracle():
    from batman_b200 import SurrogateModel, UQ
    p, n, N, seed = 8, 1024, 100000, 123456
    st = synthetic_state(n, p, 'matern32', ls=np.array([0.8, 1.2, 1.7, 2.2, 2.8, 3.4, 4.0, 4.6]), c=2.0)
    s = SurrogateModel('kriging', np.array([[0.0] * p, [1.0] * p]), None)
    s.adopt(kriging_from_states([st]), st['X_train'])
    uq = UQ(s, dists=['Uniform(0., 1.)'] * p, nsample=N, seed=seed)
    s2, s1, stot = uq.sobol()
    A, B = philox.base_samples(seed, 0, N, [('uniform', 0.0, 1.0)] * p)
    Y = _oracle_means([st], saltelli.design_from_base(A, B))
    want = saltelli.uq_sobol_aggregate(Y, N, p)
    np.testing.assert_allclose(s1, want[1], rtol=0, atol=ATOL_S)
    np.testing.assert_allclose(stot, want[2], rtol=0, atol=ATOL_S)
    np.testing.assert_allclose(s2, want[0], rtol=0, atol=ATOL_S)
    est = saltelli.saltelli_indices(Y, N, p)
    np.testing.assert_allclose(uq.details['var'], est['var'], rtol=1e-10)
    # analytic anchor of the G-function (fixed-theta surrogate error + Monte-Carlo error): block roles are right
    a1, _ = ofun.g_function_indices(p)
    assert np.max(np.abs(s1 - a1)) < 0.05


# --------------------------------------------------------------------------------------------- cfg4
def _channel_flow_pod_model(n_snap=128, n_modes=20):
    corners = np.array([[15.0, 2500.0], [60.0, 6000.0]])
    X_raw = ofun.halton(n_snap, 2, corners)
    snaps = ofun.channel_flow(X_raw, dx=100.0)                       # (n_snap, 400)
    mean_snapshot, U, S, V = opod.fit(snaps.T, 1.0 - 1e-12, n_modes)
    assert U.shape == (400, n_modes)
    modes = V * S                                                      # (n_snap, m): pod.py VS, what the GPs learn
    Xs = (X_raw - corners[0]) / (corners[1] - corners[0])
    states = []
    for q in range(n_modes):
        if q < 4:        # leading modes: smooth, well conditioned
            tree = ('prod', ('const', 1.0), ('matern', 1.5, np.array([0.6, 0.9])))
        elif q < 6:      # a fit that ran into the LOWER length-scale bound: L^-1 is near-diagonal, its rows span
            tree = ('prod', ('const', 1.0), ('matern', 1.5, np.array([0.05, 0.08])))   # more than 2^40
        elif q < 14:     # middle: Matern-5/2 with long length scales
            tree = ('prod', ('const', 0.5), ('matern', 2.5, np.array([1.5, 2.5])))
        else:            # trailing near-noise modes fitted by a wide RBF: K is singular down to the jitter,
            tree = ('prod', ('const', 1.0), ('rbf', np.array([1.2, 1.8])))   # rows of L^-1 span > 2^40
        states.append(ogp.fit_state(tree, Xs, modes[:, q]))
    return corners, X_raw, mean_snapshot, U, states


def test_cfg4_pod_field_F400_m20_predict_and_sobol():
    from batman_b200 import Pod, SurrogateModel, UQ
    corners, X_raw, mean_snapshot, U, states = _channel_flow_pod_model()
    # the two stress cases the int8 slicing (fixed point relative to each row's largest entry) has to survive
    Wa = np.abs(np.tril(np.linalg.inv(states[4]['L'])))
    row_range = max(Wa[i, :i + 1].max() / Wa[i, :i + 1][Wa[i, :i + 1] > 0].min() for i in range(len(Wa)))
    assert row_range > 2.0 ** 40, "stress case 1: a row of L^-1 spanning more than 2^40"
    assert np.linalg.cond(states[-1]['L']) > 1e5, "stress case 2: K singular down to the jitter"
    s = SurrogateModel('kriging', corners, ['Ks', 'Q'])
    s.adopt(kriging_from_states(states), X_raw, pod=Pod(mean_snapshot, U))

    # (i) SurrogateModel.__call__: modes -> field for mean and sigma (surrogate_model.py:191-196)
    rng = np.random.RandomState(7)
    pts = corners[0] + (corners[1] - corners[0]) * rng.rand(300, 2)
    mean, sigma = s(pts)
    assert mean.shape == (300, 400) and sigma.shape == (300, 400)
    scaled = (pts - corners[0]) / (corners[1] - corners[0])
    om, osig = ogp.evaluate(states, scaled)
    # mode-level gates first (1e-10 of the output scale, widened to the conditioning floor of singular K)
    km, ksig = s.predictor.evaluate(scaled)
    for q, st in enumerate(states):
        scale = np.maximum(np.abs(om[:, q]), st['y_std'])
        amp = gp_hp.amplification(st, scaled) * np.abs(om[:, q] - st['y_mean']) / scale
        gate = np.maximum(RTOL, 8 * EPS * amp)
        assert np.all(np.abs(km[:, q] - om[:, q]) / scale <= gate), (q, amp.max())
        prior = ogp.kernel_diag(st['tree']) * st['y_std'] ** 2
        # variance: k* is exact to ~1 ulp, L^-1 k* amplifies it by |W|: floor 8 eps |W| |k*| relative to the prior
        W = np.linalg.inv(st['L'])
        Kt = ogp.kernel_cross(st['tree'], scaled, st['X_train'])
        v = Kt @ W.T
        floor = 8 * EPS * 2.0 * (np.abs(v) * (np.abs(Kt) @ np.abs(W).T)).sum(axis=1) * st['y_std'] ** 2
        assert np.all(np.abs(ksig[:, q] ** 2 - osig[:, q] ** 2) <= np.maximum(RTOL * prior, floor)), q
    want_mean = opod.inverse_transform(mean_snapshot, U, om)
    want_sigma = opod.inverse_transform(mean_snapshot, U, osig)
    np.testing.assert_allclose(mean, want_mean, rtol=0, atol=1e-10 * np.abs(want_mean).max())
    np.testing.assert_allclose(sigma, want_sigma, rtol=0, atol=1e-7 * np.abs(want_sigma).max())

    # (ii) UQ.sobol(): p = 2 -> N(p+2) rows; aggregated indices and the per-feature map vs the oracle
    N, seed = 2000, 123456
    dists = ['Uniform(15., 60.)', 'Normal(4035., 400.)']
    uq = UQ(s, dists=dists, nsample=N, seed=seed)
    s2, s1, stot = uq.sobol()
    A, B = philox.base_samples(seed, 0, N, [('uniform', 15.0, 60.0), ('normal', 4035.0, 400.0)])
    D = saltelli.design_from_base(A, B)
    Y = opod.inverse_transform(mean_snapshot, U, _oracle_means(states, (D - corners[0]) / (corners[1] - corners[0])))
    want = saltelli.uq_sobol_aggregate(Y, N, 2)
    np.testing.assert_allclose(s1, want[1], rtol=0, atol=ATOL_S)
    np.testing.assert_allclose(stot, want[2], rtol=0, atol=ATOL_S)
    np.testing.assert_allclose(s2, want[0], rtol=0, atol=ATOL_S)
    with np.errstate(invalid='ignore', divide='ignore'):
        est = saltelli.saltelli_indices(Y, N, 2)
    live = est['var'] > 1e-12 * est['var'].max()      # the downstream boundary value is a constant: S = 0/0 there
    assert live.sum() >= 395
    np.testing.assert_allclose(uq.details['S1'][live], est['S1'][live], rtol=0, atol=ATOL_S)
    np.testing.assert_allclose(uq.details['ST'][live], est['ST'][live], rtol=0, atol=ATOL_S)


# --------------------------------------------------------------------------------------------- cfg5
def test_cfg5_n16384_p20_rbf_mean_sigma_both_kernels():
    p, n = 20, 16384
    X = philox.uniforms(123456, np.arange(n), p, 0)
    y = sum(np.sin(2 * np.pi * (i + 1) * X[:, i]) / (i + 1) for i in range(p))
    ls = 0.3 * (1.0 + np.arange(1, p + 1) / 20.0)
    st = ogp.fit_state(('prod', ('const', 1.0), ('rbf', ls)), X, y)
    k = kriging_from_states([st])
    rng = np.random.RandomState(5)
    pts = np.vstack([rng.rand(250, p), X[:3], X[:3] + 1e-6])
    model = k._model()
    assert model.sigma_kernel == 'int8' and model.int8_scheme == 0      # n_pad > 8192: 8 x 7-bit signed slices
    mean, sigma = k.evaluate(pts)
    omean, osigma = ogp.evaluate([st], pts)
    scale = np.maximum(np.abs(omean[:, 0]), st['y_std'])
    amp = gp_hp.amplification(st, pts) * np.abs(omean[:, 0] - st['y_mean']) / scale
    gate = np.maximum(RTOL, 8 * EPS * amp)
    assert np.all(np.abs(mean[:, 0] - omean[:, 0]) / scale <= gate), amp.max()
    prior = ogp.kernel_diag(st['tree']) * st['y_std'] ** 2
    assert np.max(np.abs(sigma[:, 0] ** 2 - osigma[:, 0] ** 2)) <= RTOL * prior
    # FP64 DMMA kernel on the same model: the two sigma kernels agree far below the gate
    model.set_sigma_kernel('dmma')
    _, sigma_dmma = k.evaluate(pts)
    assert np.max(np.abs(sigma_dmma[:, 0] ** 2 - osigma[:, 0] ** 2)) <= RTOL * prior
    assert np.max(np.abs(sigma_dmma[:, 0] ** 2 - sigma[:, 0] ** 2)) <= 1e-12 * prior
    # long-double arbiter (n forward-substitution steps on 6 points)
    sub = slice(0, 6)
    mld, sld = gp_hp.predict_hp(st, pts[sub])
    assert np.all(np.abs(mean[sub, 0] - mld.astype(float)) / scale[sub] <= gate[sub])
    assert np.max(np.abs(sigma[sub, 0] ** 2 - (sld ** 2).astype(float))) <= RTOL * prior


# --------------------------------------------------------------------------------------------- cfg3 (fit side)
def test_cfg3_lml_n4096_vs_sklearn():
    import torch
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    from batman_b200 import kernel_spec
    from batman_b200.fit import normalise_y
    from batman_b200.lml import lml_batched
    n, p = 4096, 10
    X = ofun.halton(n, p)
    y = ofun.michalewicz(X * np.pi)[:, 0]
    yn, _, _ = normalise_y(y)
    kern = ConstantKernel(1.0) * Matern(length_scale=[0.5] * p, nu=1.5)
    gp = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, y)
    rng = np.random.RandomState(2)
    thetas = [kern.theta, kern.theta + rng.uniform(-1.0, 1.0, size=p + 1)]
    want = np.array([gp.log_marginal_likelihood(t) for t in thetas])
    specs = [kernel_spec.flatten(kern.clone_with_theta(t), p) for t in thetas]
    got = lml_batched(torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda(), specs)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-6)


# --------------------------------------------------------------------------------------------- facade state
def test_predictor_evaluate_contract_survives_facade_calls():
    """surrogate(raw) / UQ.sobol() fuse the MinMax scaler for their own call only: afterwards
    ``predictor.evaluate`` still takes ALREADY-SCALED points (kriging.py:192), corners != [0, 1]."""
    import pickle
    from batman_b200 import SurrogateModel, UQ
    from conftest import load_golden
    z, states = load_golden('ishigami_n100')                           # corners = [-pi, pi]^3
    s = SurrogateModel('kriging', z['corners'], None)
    s.adopt(kriging_from_states(states), z['X_raw'])
    lo, hi = z['corners']
    scaled = (z['points'] - lo) / (hi - lo)
    before = s.predictor.evaluate(scaled)
    raw_mean, raw_sigma = s(z['points'])
    after = s.predictor.evaluate(scaled)
    np.testing.assert_array_equal(before[0], after[0])
    np.testing.assert_array_equal(before[1], after[1])
    np.testing.assert_allclose(raw_mean, before[0], rtol=0, atol=1e-12)   # raw through the facade == scaled direct
    UQ(s, dists=['Uniform(-3.14, 3.14)'] * 3, nsample=256).sobol()
    again = s.predictor.evaluate(scaled)
    np.testing.assert_array_equal(before[0], again[0])
    k2 = pickle.loads(pickle.dumps(s.predictor))                        # nothing sticky in the pickle either
    np.testing.assert_array_equal(k2.evaluate(scaled)[0], before[0])
    m3, _ = s(z['points'])                                              # and the facade still scales afterwards
    np.testing.assert_array_equal(m3, raw_mean)
