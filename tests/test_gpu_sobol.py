"""GPU parity: UQ.sobol() through the C ABI vs the oracle (GP means -> Saltelli estimator).

Gate (north_star): Sobol' indices within 1e-8 absolute of the reference path on
the same inputs and seeds.  The estimator half of the oracle is a documented
restatement of OpenTURNS 1.15 (parity with OT itself unpinned: oracle/saltelli.py).
"""
import ctypes

import numpy as np
import pytest

from conftest import kriging_from_states, load_golden, synthetic_state
from oracle import functions as ofun
from oracle import gp as ogp
from oracle import philox, pod as opod, saltelli

pytestmark = pytest.mark.gpu

ATOL = 1e-8


def _surrogate(case, pod=False):
    from batman_b200 import Pod, SurrogateModel
    z, states = load_golden(case)
    s = SurrogateModel('kriging', z['corners'], None)
    s.adopt(kriging_from_states(states), z['X_raw'], pod=Pod(z['mean_snapshot'], z['U']) if pod else None)
    return z, states, s


def _oracle_outputs(z, states, X, pod=False):
    lo, hi = z['corners']
    Y = ogp.evaluate(states, (X - lo) / (hi - lo))[0]
    if pod:
        Y = opod.inverse_transform(z['mean_snapshot'], z['U'], Y)
    return Y


def test_device_generator_is_bit_exact_with_oracle_philox():
    import torch
    from batman_b200 import _lib
    lib = _lib.load()
    p, cnt, row_lo, seed = 5, 1000, (1 << 33) + 12345, 123456     # rows beyond 2^32 exercise the high counter word
    dists = [('uniform', -3.0, 2.0), ('uniform', 0.0, 1.0), ('normal', 4035.0, 400.0), ('uniform', 15.0, 60.0),
             ('normal', 0.0, 1.0)]
    kinds = _lib.as_int_array([0 if d[0] == 'uniform' else 1 for d in dists])
    params = _lib.as_double_array([v for d in dists for v in list(d[1:]) + [0.0, 0.0]])     # 4 parameters per marginal
    for stream in (0, 1):
        out = torch.empty((cnt, p), dtype=torch.float64, device='cuda')
        _lib.check(lib.bk_design_rows(seed, row_lo, cnt, p, kinds, params, stream, ctypes.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        want = philox.marginal_transform(philox.uniforms(seed, np.arange(row_lo, row_lo + cnt), p, stream), dists)
        for j, d in enumerate(dists):
            if d[0] == 'uniform':
                assert np.array_equal(got[:, j], want[:, j])
            else:
                np.testing.assert_allclose(got[:, j], want[:, j], rtol=1e-13, atol=1e-13 * d[2])


@pytest.mark.parametrize('N', [1, 37, 2000])
def test_sobol_ishigami_model_vs_oracle_same_design(N):
    """cfg1 model (Ishigami, n_train=100): same base samples, estimator parity 1e-8; ragged N."""
    from batman_b200 import UQ
    z, states, s = _surrogate('ishigami_n100')
    if N < 2:
        with pytest.raises(Exception):
            UQ(s, dists=['Uniform(-3.14, 3.14)'] * 3, nsample=N).sobol()
        return
    dists = [('uniform', -np.pi, np.pi)] * 3
    A, B = philox.base_samples(123456, 0, N, dists)
    uq = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3, nsample=N, base_samples=(A, B))
    s2, s1, st = uq.sobol()
    Y = _oracle_outputs(z, states, saltelli.design_from_base(A, B))
    want = saltelli.uq_sobol_aggregate(Y, N, 3)
    est = saltelli.saltelli_indices(Y, N, 3)
    np.testing.assert_allclose(s1, want[1], rtol=0, atol=ATOL)
    np.testing.assert_allclose(st, want[2], rtol=0, atol=ATOL)
    np.testing.assert_allclose(s2, want[0], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uq.details['S2'][0], est['S2'][0], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uq.details['var'], est['var'], rtol=1e-10)
    jan = saltelli.jansen_indices(Y, N, 3)
    np.testing.assert_allclose(uq.details['S1_jansen'], jan['S1'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uq.details['ST_jansen'], jan['ST'], rtol=0, atol=ATOL)
    # in-kernel generator == explicit samples (the uniform marginal is bit-exact)
    uq2 = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3, nsample=N, seed=123456)
    g2, g1, gt = uq2.sobol()
    np.testing.assert_allclose(g1, s1, rtol=0, atol=1e-13)
    np.testing.assert_allclose(gt, st, rtol=0, atol=1e-13)
    np.testing.assert_allclose(g2, s2, rtol=0, atol=1e-13)


def test_sobol_ishigami_matches_analytic_like_the_reference_test():
    """tests/test_uq.py:9-29 of the reference: |dS| <~ 0.05-0.1 against the closed forms."""
    from batman_b200 import UQ
    z, states, s = _surrogate('ishigami_n100')
    s2, s1, st = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3, nsample=20000).sobol()
    a1, at, a2 = ofun.ishigami_indices()
    assert np.sqrt(np.sum((a1 - s1) ** 2)) < 0.1
    assert np.sqrt(np.sum((at - st) ** 2)) < 0.1
    assert np.sqrt(np.sum((a2 - s2) ** 2)) < 0.3


def test_sobol_multi_output_aggregated():
    """3 independent GPs (m = 3): per-output and aggregated indices, aggregated S2 weights (uq.py:403-420)."""
    from batman_b200 import SurrogateModel, UQ
    z, states = load_golden('kernel_default_3out')
    corners = np.array([[0.0] * 3, [1.0] * 3])
    s = SurrogateModel('kriging', corners, None)
    s.adopt(kriging_from_states(states), z['X_train'])
    N = 1500
    A, B = philox.base_samples(7, 0, N, [('uniform', 0.0, 1.0)] * 3)
    uq = UQ(s, dists=['Uniform(0., 1.)'] * 3, nsample=N, base_samples=(A, B))
    s2, s1, st = uq.sobol()
    Y = ogp.evaluate(states, saltelli.design_from_base(A, B))[0]
    want = saltelli.uq_sobol_aggregate(Y, N, 3)
    est = saltelli.saltelli_indices(Y, N, 3)
    np.testing.assert_allclose(uq.details['S1'], est['S1'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uq.details['ST'], est['ST'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(uq.details['S2'], est['S2'], rtol=0, atol=ATOL)
    for got, w in zip((s2, s1, st), want):
        np.testing.assert_allclose(got, w, rtol=0, atol=ATOL)


@pytest.mark.parametrize('indices', ['aggregated', 'block'])
def test_sobol_pod_field_p2(indices):
    """Channel_Flow POD surrogate (p = 2 -> N(p+2) rows; Normal marginal via explicit samples)."""
    from batman_b200 import UQ
    z, states, s = _surrogate('channel_flow_pod', pod=True)
    N = 3000
    dists = [('uniform', 15.0, 60.0), ('normal', 4035.0, 400.0)]
    A, B = philox.base_samples(123456, 0, N, dists)
    uq = UQ(s, dists=['Uniform(15., 60.)', 'Normal(4035., 400.)'], nsample=N, indices=indices, base_samples=(A, B))
    got = uq.sobol()
    X = saltelli.design_from_base(A, B)
    assert len(X) == 4 * N
    Y = _oracle_outputs(z, states, X, pod=True)
    if indices == 'block':
        Y = saltelli.int_func(Y)
    want = saltelli.uq_sobol_aggregate(Y, N, 2, indices=indices)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=0, atol=ATOL)
    # generator path: Normal marginal differs from scipy's ndtri by ulps only
    got2 = UQ(s, dists=['Uniform(15., 60.)', 'Normal(4035., 400.)'], nsample=N, indices=indices).sobol()
    for g, w in zip(got2, want):
        np.testing.assert_allclose(g, w, rtol=0, atol=1e-7)


def test_sobol_ensemble_mode_user_supplied_outputs():
    """uq.py:336-339: no surrogate, Saltelli design + outputs given."""
    from batman_b200 import UQ
    N, p = 4000, 3
    A, B = philox.base_samples(3, 0, N, [('uniform', -np.pi, np.pi)] * p)
    X = saltelli.design_from_base(A, B)
    Y = np.hstack([ofun.ishigami(X), 2.0 * ofun.ishigami(X, a=3.0) + 5.0])
    uq = UQ(None, dists=['Uniform(-3.14, 3.14)'] * p, space=X, data=Y)
    got = uq.sobol()
    want = saltelli.uq_sobol_aggregate(Y, N, p)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=0, atol=ATOL)
    est = saltelli.saltelli_indices(Y, N, p)
    np.testing.assert_allclose(uq.details['S1'], est['S1'], rtol=0, atol=ATOL)


def test_sobol_partials_are_additive_over_row_ranges():
    """What multi-GPU sharding relies on: sums over [0,N) == sums over [0,k) + [k,N)."""
    import torch
    from batman_b200 import _lib
    z, states, s = _surrogate('ishigami_n100')
    lib = _lib.load()
    model = s.predictor._model()
    model.set_scaler(s.scaler.scale_, s.scaler.min_)      # low-level handle state; the facades set it per call
    p, N = 3, 1000
    nsums, nslots = lib.bk_sobol_nsums(p, 1), lib.bk_sobol_nslots()
    kinds = _lib.as_int_array([0] * p)
    params = _lib.as_double_array([-np.pi, np.pi, 0.0, 0.0] * p)
    shift = _lib.as_double_array(model.y_means)

    def totals(ranges):
        part = torch.zeros((nslots, 1, nsums, 2), dtype=torch.float64, device='cuda')
        for lo, hi in ranges:
            _lib.check(lib.bk_sobol_accumulate(model.handle, None, None, 99, kinds, params, lo, hi - lo, 1, shift,
                                               ctypes.c_void_p(part.data_ptr()), None))
        tot = torch.empty((1, nsums, 2), dtype=torch.float64, device='cuda')
        _lib.check(lib.bk_sobol_reduce_slots(ctypes.c_void_p(part.data_ptr()), 1, nsums, ctypes.c_void_p(tot.data_ptr()), None))
        torch.cuda.synchronize()
        t = tot.cpu().numpy()
        return t[..., 0] + t[..., 1]

    full = totals([(0, N)])
    split = totals([(0, 333), (333, 800), (800, N)])
    np.testing.assert_allclose(split, full, rtol=1e-13, atol=1e-13)


def test_sobol_full_size_cfg2_properties():
    """BASELINE config 2 at full size (G-function 8-D, n_train=1024, N=1e6, R=1.8e7 predictions).
    Size-independent properties: analytic anchor (surrogate + MC error), Saltelli ~ Jansen,
    first <= total, and the sums' additivity (above)."""
    from batman_b200 import SurrogateModel, UQ
    st = synthetic_state(1024, 8, 'matern32', ls=np.full(8, 1.5))
    s = SurrogateModel('kriging', np.array([[0.0] * 8, [1.0] * 8]), None)
    s.adopt(kriging_from_states([st]), st['X_train'])
    uq = UQ(s, dists=['Uniform(0., 1.)'] * 8, nsample=1000000)
    s2, s1, st_ = uq.sobol()
    a1, _ = ofun.g_function_indices(8)
    at = ofun.g_function_total_exact(8)
    assert np.max(np.abs(s1 - a1)) < 0.05 and np.max(np.abs(st_ - at)) < 0.05
    assert np.max(np.abs(uq.details['S1_jansen'][0] - s1)) < 0.01
    assert np.max(np.abs(uq.details['ST_jansen'][0] - st_)) < 0.01
    assert np.all(st_ > s1 - 0.01)
    assert s2.shape == (8, 8) and np.allclose(s2, s2.T) and np.all(np.diag(s2) == 0)


def test_sobol_block_indices_without_pod():
    """indices='block' on a 3-output Kriging: trapz over the outputs (uq.py:205-218, 319-324)."""
    from batman_b200 import SurrogateModel, UQ
    z, states = load_golden('kernel_default_3out')
    s = SurrogateModel('kriging', np.array([[0.0] * 3, [1.0] * 3]), None)
    s.adopt(kriging_from_states(states), z['X_train'])
    N = 1200
    A, B = philox.base_samples(11, 0, N, [('uniform', 0.0, 1.0)] * 3)
    got = UQ(s, dists=['Uniform(0., 1.)'] * 3, nsample=N, indices='block', base_samples=(A, B)).sobol()
    Y = saltelli.int_func(ogp.evaluate(states, saltelli.design_from_base(A, B))[0])
    want = saltelli.uq_sobol_aggregate(Y, N, 3, indices='block')
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=0, atol=ATOL)


@pytest.mark.parametrize('k,m,F', [(1, 3, 3), (1000, 20, 400), (77, 5, 7), (4096, 64, 33)])
def test_pod_inverse_dmma_kernel(k, m, F):
    """Pod.inverse_transform on the FP64 tensor cores vs numpy (pod/pod.py:226-228)."""
    import torch
    from batman_b200.pod_kernels import pod_inverse
    rng = np.random.RandomState(k)
    U, mean, modes = rng.randn(F, m), rng.randn(F) * 10, rng.randn(k, m) * 3
    got = pod_inverse(torch.from_numpy(mean).cuda(), torch.from_numpy(U).cuda(), torch.from_numpy(modes).cuda()).cpu().numpy()
    want = opod.inverse_transform(mean, U, modes)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max())


def test_pod_inverse_reference_known_answer():
    """tests/test_pod.py:109-114 of the reference through the device kernel."""
    from batman_b200 import Pod
    z, _ = load_golden('pod_inverse')
    out = Pod(z['mean_snapshot'], z['U']).inverse_transform(z['modes'])
    np.testing.assert_almost_equal(out, [[40, 43, 41], [41, 47, 42]], decimal=3)
    np.testing.assert_allclose(out, z['ref_inverse'], rtol=0, atol=1e-12)


def test_sobol_first_and_total_only():
    """second_order=False: the design has no C blocks (N(p+2) predictions), centring over the rows that exist;
    first / total order and Jansen against the oracle on the same base samples; the fused and the POD paths."""
    from batman_b200 import UQ
    z, states, s = _surrogate('ishigami_n100')
    N = 3000
    dists = [('uniform', -np.pi, np.pi)] * 3
    A, B = philox.base_samples(123456, 0, N, dists)
    uq = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3, nsample=N, base_samples=(A, B),
            second_order=False)
    s2, s1, st = uq.sobol()
    Y = _oracle_outputs(z, states, saltelli.design_from_base(A, B, second_order=False))
    assert Y.shape[0] == 5 * N
    est = saltelli.saltelli_indices(Y, N, 3, second_order=False)
    np.testing.assert_allclose(s1, est['S1_agg'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(st, est['ST_agg'], rtol=0, atol=ATOL)
    assert np.all(s2 == 0.0)
    jan = saltelli.jansen_indices(Y, N, 3)
    np.testing.assert_allclose(uq.details['ST_jansen'], jan['ST'], rtol=0, atol=ATOL)
    # ensemble mode on the stored p+2 blocks gives the same numbers
    ens = UQ(None, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3,
             space=saltelli.design_from_base(A, B, second_order=False), data=Y, second_order=False)
    e2, e1, et = ens.sobol()
    np.testing.assert_allclose(e1, s1, rtol=0, atol=ATOL)
    np.testing.assert_allclose(et, st, rtol=0, atol=ATOL)
    # POD field path
    zp, statesp, sp = _surrogate('channel_flow_pod', pod=True)
    Np = 500
    Ap, Bp = philox.base_samples(7, 0, Np, [('uniform', 15.0, 60.0), ('normal', 4035.0, 400.0)])
    g2, g1, gt = UQ(sp, dists=['Uniform(15., 60.)', 'Normal(4035., 400.)'], nsample=Np, base_samples=(Ap, Bp),
                    second_order=False).sobol()
    Yp = _oracle_outputs(zp, statesp, saltelli.design_from_base(Ap, Bp, second_order=False), pod=True)
    estp = saltelli.saltelli_indices(Yp, Np, 2, second_order=False)
    np.testing.assert_allclose(g1, estp['S1_agg'], rtol=0, atol=ATOL)
    np.testing.assert_allclose(gt, estp['ST_agg'], rtol=0, atol=ATOL)


def test_sobol_intervals_jackknife_vs_oracle_and_files(tmp_path):
    """95 % intervals of the aggregated indices: device leave-one-group-out estimates (bk_sobol_jackknife over the
    partial slots; slot of base row k in the fused kernel = (k // 128) % 592) against the numpy jackknife on the same
    groups; sensitivity.json / sensitivity_aggregated.json in the reference's layout (uq.py:380-393, 437-463)."""
    import json
    from batman_b200 import UQ
    z, states, s = _surrogate('ishigami_n100')
    N = 20000
    dists = [('uniform', -np.pi, np.pi)] * 3
    A, B = philox.base_samples(123456, 0, N, dists)
    uq = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3, nsample=N, base_samples=(A, B),
            fname=str(tmp_path), plabels=['x1', 'x2', 'x3'], intervals='jackknife')
    s2, s1, st = uq.sobol()
    d = uq.details
    assert d['interval_method'] == 'jackknife'
    assert d['jackknife_groups'] == (N + 127) // 128
    Y = _oracle_outputs(z, states, saltelli.design_from_base(A, B))
    theta, sigma = saltelli.jackknife_aggregated(Y, N, 3, (np.arange(N) // 128) % 592)
    np.testing.assert_allclose(np.concatenate([s1, st]), theta, rtol=0, atol=ATOL)
    np.testing.assert_allclose(np.concatenate([d['S1_agg_sigma'], d['ST_agg_sigma']]), sigma, rtol=1e-6, atol=1e-10)
    lo, hi = d['S1_interval']
    assert np.all(lo < s1) and np.all(s1 < hi) and np.all(hi - lo < 0.2)
    a1, at, _ = ofun.ishigami_indices()
    agg = json.load(open(tmp_path / 'sensitivity_aggregated.json'))
    assert list(agg) == [a + lab for a in ['S_min_', 'S_', 'S_max_', 'S_T_min_', 'S_T_', 'S_T_max_'] for lab in ['x1', 'x2', 'x3']]
    assert agg['S_x1'] == [[s1[0]]] and agg['S_T_max_x3'] == [[d['ST_interval'][1][2]]]
    sens = json.load(open(tmp_path / 'sensitivity.json'))
    assert list(sens) == ['S_x1', 'S_x2', 'S_x3', 'S_T_x1', 'S_T_x2', 'S_T_x3'] and sens['S_T_x2'] == [[d['ST'][0, 1]]]
    # POD field: F = 400 features share the groups; intervals exist and the map file has the x column
    zp, statesp, sp = _surrogate('channel_flow_pod', pod=True)
    uqp = UQ(sp, dists=['Uniform(15., 60.)', 'Normal(4035., 400.)'], nsample=3000, fname=str(tmp_path / 'pod'))
    uqp.sobol()
    assert np.all(np.isfinite(uqp.details['S1_agg_sigma'])) and uqp.details['jackknife_groups'] > 100
    m = json.load(open(tmp_path / 'pod' / 'sensitivity.json'))
    assert list(m)[0] == 'x' and len(m['S_x0']) == sp.pod.mean_snapshot.shape[0]
