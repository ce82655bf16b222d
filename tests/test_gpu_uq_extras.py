"""GPU parity of the steps around the Saltelli pass:

* device inverse CDFs of every supported marginal vs scipy.special on the same Philox uniforms;
* the LHS propagation sample, its device-side moments (uq.py:159-173, 501-529) and ``error_model`` (uq.py:220-284);
* the other OpenTURNS estimators uq.py:340 names (Jansen, Martinez, Mauntz-Kucherenko) from the same sums;
* the asymptotic (delta-method) intervals of uq.py:341, 425-426 vs the oracle restatement -- single output, several
  outputs, POD field, block integral, ensemble mode.
The OpenTURNS half of the oracle is a documented restatement (parity with OT itself unpinned: oracle/saltelli.py).
"""
import ctypes
import json

import numpy as np
import pytest

from conftest import kriging_from_states, load_golden, synthetic_state
from oracle import functions as ofun
from oracle import gp as ogp
from oracle import philox, pod as opod, saltelli

pytestmark = pytest.mark.gpu

ATOL = 1e-8

ALL_DISTS = [
    ('Uniform(-3., 2.)', ('uniform', -3.0, 2.0)),
    ('Normal(4035., 400.)', ('normal', 4035.0, 400.0)),
    ('LogNormal(0.3, 0.6, 1.5)', ('lognormal', 0.3, 0.6, 1.5)),
    ('Triangular(1., 2., 5.)', ('triangular', 1.0, 2.0, 5.0)),
    ('TruncatedNormal(1., 2., -1., 2.5)', None),
    ('Beta(2.5, 4., -1., 3.)', ('beta', 2.5, 4.0, -1.0, 3.0)),
    ('BetaMuSigma(4031, 400, 1000, 6000).getDistribution()', None),
    ('Beta(0.6, 0.8, 0., 1.)', ('beta', 0.6, 0.8, 0.0, 1.0)),
    ('Gumbel(2., 1.)', ('gumbel', 2.0, 1.0)),
    ('Exponential(0.5, 2.)', ('exponential', 0.5, 2.0)),
    ('LogUniform(-1., 2.)', ('loguniform', -1.0, 2.0)),
    ('WeibullMin(2., 1.5, 0.5)', ('weibullmin', 2.0, 1.5, 0.5)),
    ('Logistic(1., 2.)', ('logistic', 1.0, 2.0)),
    ('Rayleigh(2., 1.)', ('rayleigh', 2.0, 1.0)),
]


def _oracle_dists(texts):
    """parsed product marginal -> oracle tuple (name + the same 4 parameters)."""
    from batman_b200 import distributions
    out = []
    for t in texts:
        d = distributions.parse(t)
        out.append(tuple([distributions.NAMES[d[0]]] + list(d[1:])))
    return out


def test_device_inverse_cdfs_match_scipy_on_the_same_uniforms():
    import torch
    from batman_b200 import _lib, distributions
    lib = _lib.load()
    texts = [t for t, _ in ALL_DISTS]
    p, cnt, seed = len(texts), 20000, 2024
    kinds, params = distributions.parse_all(texts)
    odists = _oracle_dists(texts)
    for (t, explicit), od in zip(ALL_DISTS, odists):           # the parser and the hand-written oracle tuples agree
        if explicit is not None:
            assert od[0] == explicit[0] and list(od[1:1 + len(explicit) - 1]) == list(explicit[1:]), t
    out = torch.empty((cnt, p), dtype=torch.float64, device='cuda')
    _lib.check(lib.bk_design_rows(seed, 5, cnt, p, _lib.as_int_array(kinds), _lib.as_double_array(params), 0,
                                  ctypes.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = philox.marginal_transform(philox.uniforms(seed, np.arange(5, 5 + cnt), p, 0), odists)
    for j, t in enumerate(texts):
        scale = np.maximum(np.abs(want[:, j]), np.std(want[:, j]))
        tol = 2e-10 if 'Beta' in t else 1e-12
        assert np.max(np.abs(got[:, j] - want[:, j]) / scale) <= tol, (t, np.max(np.abs(got[:, j] - want[:, j]) / scale))
    with pytest.raises(_lib.BatmanB200Error):
        lib_kinds = _lib.as_int_array([99] * p)
        _lib.check(lib.bk_design_rows(seed, 0, cnt, p, lib_kinds, _lib.as_double_array(params), 0,
                                      ctypes.c_void_p(out.data_ptr()), None))


def _ishigami_surrogate():
    from batman_b200 import SurrogateModel
    z, states = load_golden('ishigami_n100')
    s = SurrogateModel('kriging', z['corners'], None)
    s.adopt(kriging_from_states(states), z['X_raw'])
    return z, states, s


def test_lhs_sample_moments_and_error_model(tmp_path):
    from batman_b200 import UQ
    z, states, s = _ishigami_surrogate()
    N = 5000
    dists = ['Uniform(-3.141592653589793, 3.141592653589793)', 'Normal(0., 1.)', 'Triangular(-3., 0., 3.)']
    uq = UQ(s, dists=dists, nsample=N, seed=7, fname=str(tmp_path), test='Ishigami')
    X = uq.sample
    assert X.shape == (N, 3)
    want = philox.lhs_sample(7, uq._lhs_perm, _oracle_dists(dists))
    np.testing.assert_allclose(X, want, rtol=1e-12, atol=1e-12)
    # LHS property: every one of the N equiprobable cells of a dimension holds exactly one point
    u0 = (X[:, 0] + np.pi) / (2 * np.pi)
    assert np.array_equal(np.sort(np.floor(u0 * N).astype(int)), np.arange(N))
    # moments on the device == numpy on the downloaded outputs == the oracle's outputs on the same sample
    mean, sd = uq.moments()
    lo, hi = z['corners']
    Y = ogp.evaluate(states, (X - lo) / (hi - lo))[0]
    np.testing.assert_allclose(uq.output, Y, rtol=0, atol=1e-10 * states[0]['y_std'])
    np.testing.assert_allclose(mean, Y.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(sd, Y.std(axis=0), rtol=1e-11)
    res = uq.error_propagation()
    np.testing.assert_allclose(res['min'], Y.min(axis=0), rtol=1e-12)
    np.testing.assert_allclose(res['max'], Y.max(axis=0), rtol=1e-12)
    assert set(json.load(open(tmp_path / 'moment.json'))) == {'min', 'sd_min', 'mean', 'sd_max', 'max'}
    # error_model (uq.py:220-284) on a uniform propagation sample, through UQ(test=...) as the reference wires it
    uq2 = UQ(s, dists=[dists[0]] * 3, nsample=2000, seed=3, fname=str(tmp_path), test='Ishigami')
    ind = uq2.sobol()
    q2, mse, l2_2nd, l2_1st, l2_tot = uq2.model_error
    from sklearn.metrics import r2_score
    Xs = uq2.sample
    Ys = ogp.evaluate(states, (Xs - lo) / (hi - lo))[0]
    assert q2 == pytest.approx(r2_score(ofun.ishigami(Xs), Ys), abs=1e-9)
    a1, at, a2 = ofun.ishigami_indices()
    assert l2_1st == pytest.approx(np.sqrt(((a1 - ind[1]) ** 2).sum()), abs=1e-12)
    fields = open(tmp_path / 'model_err.dat').read().split()
    assert len(fields) == 6 and float(fields[1]) == pytest.approx(q2)
    with pytest.raises(NotImplementedError):
        UQ(s, dists=[dists[0]] * 3, nsample=64, test='Forrester').sobol()


def _three_output_model():
    from batman_b200 import SurrogateModel
    X = ofun.halton(160, 3, [[-np.pi] * 3, [np.pi] * 3])
    Xs = (X + np.pi) / (2 * np.pi)
    ys = [ofun.ishigami(X)[:, 0], ofun.ishigami(X, a=3.0)[:, 0] * 2.0 + 5.0, np.cos(X).sum(axis=1) * 0.3 - 2.0]
    trees = [('prod', ('const', 1.0), ('matern', 1.5, np.array([0.4, 0.5, 0.6]))),
             ('prod', ('const', 2.0), ('matern', 2.5, np.array([0.5, 0.5, 0.5]))),
             ('prod', ('const', 0.7), ('rbf', np.array([0.6, 0.7, 0.8])))]
    states = [ogp.fit_state(t, Xs, y) for t, y in zip(trees, ys)]
    s = SurrogateModel('kriging', np.array([[-np.pi] * 3, [np.pi] * 3]), None)
    s.adopt(kriging_from_states(states), X)
    return states, s


def _outputs(states, D):
    return ogp.evaluate(states, (D + np.pi) / (2 * np.pi))[0]


@pytest.mark.parametrize('estimator', ['jansen', 'martinez', 'mauntz-kucherenko'])
def test_other_estimators_from_the_same_sums(estimator):
    from batman_b200 import UQ
    states, s = _three_output_model()
    N, p = 3000, 3
    A, B = philox.base_samples(11, 0, N, [('uniform', -np.pi, np.pi)] * p)
    uq = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * p, nsample=N, seed=11, estimator=estimator)
    s2, s1, st = uq.sobol()
    Y = _outputs(states, saltelli.design_from_base(A, B))
    fn = {'jansen': saltelli.jansen_indices, 'martinez': saltelli.martinez_indices,
          'mauntz-kucherenko': saltelli.mauntz_kucherenko_indices}[estimator]
    want = fn(Y, N, p)
    np.testing.assert_allclose(s1, want['S1_agg'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(st, want['ST_agg'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uq.details['S1_selected'], want['S1'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uq.details['ST_selected'], want['ST'], rtol=0, atol=ATOL)
    # the Saltelli numbers of the same pass are untouched, second order is always Saltelli's
    ref = saltelli.uq_sobol_aggregate(Y, N, p)
    np.testing.assert_allclose(uq.details['S1_agg'], ref[1], rtol=0, atol=ATOL)
    np.testing.assert_allclose(s2, ref[0], rtol=0, atol=ATOL)


def test_asymptotic_intervals_multi_output_and_files(tmp_path):
    from batman_b200 import UQ
    states, s = _three_output_model()
    N, p = 4000, 3
    A, B = philox.base_samples(5, 0, N, [('uniform', -np.pi, np.pi)] * p)
    uq = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * p, nsample=N, seed=5, fname=str(tmp_path),
            plabels=['a', 'b', 'c'])
    s2, s1, st = uq.sobol()
    d = uq.details
    assert d['interval_method'] == 'asymptotic'
    Y = _outputs(states, saltelli.design_from_base(A, B))
    w1, wt = saltelli.asymptotic_sigma_aggregated(Y, N, p)
    np.testing.assert_allclose(d['S1_agg_sigma'], w1, rtol=1e-7)
    np.testing.assert_allclose(d['ST_agg_sigma'], wt, rtol=1e-7)
    np.testing.assert_allclose(d['mean'], Y.mean(axis=0), rtol=1e-10)
    # the two interval estimators answer the same question
    assert np.all(d['S1_agg_sigma_jackknife'] < 3 * w1 + 1e-3) and np.all(w1 < 3 * d['S1_agg_sigma_jackknife'] + 1e-3)
    agg = json.load(open(tmp_path / 'sensitivity_aggregated.json'))
    assert agg['S_min_a'] == [[s1[0] - 1.959963984540054 * d['S1_agg_sigma'][0]]]
    assert agg['S_T_max_c'] == [[st[2] + 1.959963984540054 * d['ST_agg_sigma'][2]]]
    # single output, explicit base samples
    z, st1, s1m = _ishigami_surrogate()
    A1, B1 = philox.base_samples(9, 0, 2500, [('uniform', -np.pi, np.pi)] * 3)
    u1 = UQ(s1m, dists=['Uniform(-3.14, 3.14)'] * 3, nsample=2500, base_samples=(A1, B1))
    u1.sobol()
    lo, hi = z['corners']
    Y1 = ogp.evaluate(st1, (saltelli.design_from_base(A1, B1) - lo) / (hi - lo))[0]
    w1, wt = saltelli.asymptotic_sigma_aggregated(Y1, 2500, 3)
    np.testing.assert_allclose(u1.details['S1_agg_sigma'], w1, rtol=1e-7)
    np.testing.assert_allclose(u1.details['ST_agg_sigma'], wt, rtol=1e-7)
    # ensemble mode: the same numbers from the stored outputs
    ue = UQ(None, dists=['Uniform(-3.14, 3.14)'] * 3, space=saltelli.design_from_base(A1, B1), data=Y1)
    ue.sobol()
    np.testing.assert_allclose(ue.details['S1_agg_sigma'], w1, rtol=1e-7)


@pytest.mark.parametrize('indices', ['aggregated', 'block'])
def test_asymptotic_intervals_pod_field_and_beta_marginal(indices):
    """Channel-flow POD surrogate with the marginals the shipped settings use (Uniform + BetaMuSigma)."""
    from batman_b200 import Pod, SurrogateModel, UQ
    z, states = load_golden('channel_flow_pod')
    s = SurrogateModel('kriging', z['corners'], None)
    s.adopt(kriging_from_states(states), z['X_raw'], pod=Pod(z['mean_snapshot'], z['U']))
    dists = ['Uniform(15., 60.)', 'BetaMuSigma(4035, 400, 2500, 6000).getDistribution()']
    N = 3000
    uq = UQ(s, dists=dists, nsample=N, seed=21, indices=indices)
    s2, s1, st = uq.sobol()
    A, B = philox.base_samples(21, 0, N, _oracle_dists(dists))
    lo, hi = z['corners']
    Ym = ogp.evaluate(states, (saltelli.design_from_base(A, B) - lo) / (hi - lo))[0]
    Y = opod.inverse_transform(z['mean_snapshot'], z['U'], Ym)
    if indices == 'block':
        Y = saltelli.int_func(Y)
    want = saltelli.uq_sobol_aggregate(Y, N, 2, indices=indices)
    # the beta inverse CDF is 1e-10 accurate, the indices inherit that
    np.testing.assert_allclose(s1, want[1], rtol=0, atol=1e-7)
    np.testing.assert_allclose(st, want[2], rtol=0, atol=1e-7)
    w1, wt = saltelli.asymptotic_sigma_aggregated(Y, N, 2)
    assert uq.details['interval_method'] == 'asymptotic'
    np.testing.assert_allclose(uq.details['S1_agg_sigma'], w1, rtol=1e-5)
    np.testing.assert_allclose(uq.details['ST_agg_sigma'], wt, rtol=1e-5)


def test_block_indices_with_more_than_64_outputs():
    """'block' indices of a Kriging model without POD integrate its m outputs: m > 64 used to hit the mode limit of the
    POD kernels (the reference has no such limit)."""
    from batman_b200 import SurrogateModel, UQ
    m, n, p = 70, 40, 2
    X = ofun.halton(n, p)
    Yt = np.stack([np.sin(3 * X[:, 0] + 0.05 * q) + (1 + 0.01 * q) * X[:, 1] ** 2 for q in range(m)], axis=1)
    states = [ogp.fit_state(('prod', ('const', 1.0), ('matern', 1.5, np.array([0.5, 0.7]))), X, Yt[:, q]) for q in range(m)]
    s = SurrogateModel('kriging', np.array([[0.0, 0.0], [1.0, 1.0]]), None)
    s.adopt(kriging_from_states(states), X)
    N = 1500
    A, B = philox.base_samples(3, 0, N, [('uniform', 0.0, 1.0)] * p)
    got = UQ(s, dists=['Uniform(0., 1.)'] * p, nsample=N, seed=3, indices='block').sobol()
    Y = saltelli.int_func(ogp.evaluate(states, saltelli.design_from_base(A, B))[0])
    want = saltelli.uq_sobol_aggregate(Y, N, p, indices='block')
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=0, atol=ATOL)
