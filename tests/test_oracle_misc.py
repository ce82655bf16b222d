"""CPU: pod / functions golden vectors, Philox known answers, Saltelli analytic anchors."""
import numpy as np

from conftest import load_golden
from oracle import functions as ofun
from oracle import philox, pod as opod, saltelli


def test_pod_inverse_known_answer():
    """tests/test_pod.py:109-114 of the reference."""
    z, _ = load_golden('pod_inverse')
    out = opod.inverse_transform(z['mean_snapshot'], z['U'], z['modes'])
    np.testing.assert_almost_equal(out, [[40, 43, 41], [41, 47, 42]], decimal=3)
    np.testing.assert_allclose(out, z['ref_inverse'], rtol=0, atol=1e-13)
    mean, U, S, V = opod.fit(z['snapshots'], 1, 3)
    np.testing.assert_almost_equal(mean, [43.5, 44.5, 43.75], decimal=1)      # test_pod.py:27
    np.testing.assert_almost_equal(S, [14.003, 4.747, 2.208], decimal=3)      # test_pod.py:28
    np.testing.assert_allclose(np.abs(U), np.abs(z['U']), atol=1e-12)


def test_functions_match_reference():
    z, _ = load_golden('functions')
    np.testing.assert_allclose(ofun.ishigami(z['x3']), z['ishigami'], rtol=0, atol=1e-13)
    np.testing.assert_allclose(ofun.g_function(z['x8']), z['g_function'], rtol=0, atol=1e-15)
    np.testing.assert_allclose(ofun.michalewicz(z['x10']), z['michalewicz'], rtol=0, atol=1e-14)
    np.testing.assert_allclose(ofun.channel_flow(z['x2'], dx=100.0), z['channel_flow'], rtol=0, atol=1e-10)
    s1, st = ofun.g_function_indices(8)
    np.testing.assert_allclose(s1, z['g_s_first'], rtol=1e-14)
    np.testing.assert_allclose(st, z['g_s_total'], rtol=1e-14)
    i1, it, i2 = ofun.ishigami_indices()
    np.testing.assert_allclose(i1, z['ishigami_s_first'], rtol=1e-14)
    np.testing.assert_allclose(it, z['ishigami_s_total2'], rtol=1e-14)
    np.testing.assert_allclose(i2, z['ishigami_s_second'], rtol=1e-14, atol=1e-16)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_philox_uniforms_open_interval_and_moments():
    u = philox.uniforms(123456, np.arange(200000), 3, 0)
    assert u.min() > 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 1e-3
    ub = philox.uniforms(123456, np.arange(200000), 3, 1)
    assert abs(np.corrcoef(u[:, 0], ub[:, 0])[0, 1]) < 1e-2
    # slabs are a pure function of the absolute row index (multi-GPU sharding relies on it)
    np.testing.assert_array_equal(philox.uniforms(123456, np.arange(1000, 2000), 3, 0), u[1000:2000])


def test_design_layout():
    A = np.arange(12.0).reshape(4, 3)
    B = -np.arange(12.0).reshape(4, 3) - 1
    D = saltelli.design_from_base(A, B, second_order=True)
    assert D.shape == (4 * 8, 3)
    np.testing.assert_array_equal(D[:4], A)
    np.testing.assert_array_equal(D[4:8], B)
    E1 = D[8 + 4:8 + 8]
    np.testing.assert_array_equal(E1[:, 1], B[:, 1])
    np.testing.assert_array_equal(E1[:, [0, 2]], A[:, [0, 2]])
    C0 = D[(2 + 3) * 4:(3 + 3) * 4]
    np.testing.assert_array_equal(C0[:, 0], A[:, 0])
    np.testing.assert_array_equal(C0[:, 1:], B[:, 1:])
    assert saltelli.design_from_base(A[:, :2], B[:, :2]).shape == (4 * 4, 2)   # p == 2: N(p+2)
    assert saltelli.design_from_base(A, B, second_order=False).shape == (4 * 5, 3)


def test_saltelli_ishigami_analytic_anchor():
    """The reference's own acceptance level is |dS| <~ 0.05-0.1 (tests/test_uq.py:23-29)."""
    N = 200000
    dists = [('uniform', -np.pi, np.pi)] * 3
    A, B = philox.base_samples(123456, 0, N, dists)
    Y = ofun.ishigami(saltelli.design_from_base(A, B))
    est = saltelli.saltelli_indices(Y, N, 3)
    s1, st, s2 = ofun.ishigami_indices()
    np.testing.assert_allclose(est['S1'][0], s1, atol=0.01)
    np.testing.assert_allclose(est['ST'][0], st, atol=0.01)
    np.testing.assert_allclose(est['S2'][0], s2, atol=0.015)
    jan = saltelli.jansen_indices(Y, N, 3)
    np.testing.assert_allclose(jan['S1'][0], s1, atol=0.01)
    np.testing.assert_allclose(jan['ST'][0], st, atol=0.01)
    agg = saltelli.uq_sobol_aggregate(Y, N, 3)
    np.testing.assert_allclose(agg[1], est['S1'][0], atol=1e-14)   # single output: aggregated == per-output


def test_saltelli_g_function_analytic_anchor():
    N, d = 100000, 8
    A, B = philox.base_samples(99, 0, N, [('uniform', 0.0, 1.0)] * d)
    Y = ofun.g_function(saltelli.design_from_base(A, B, second_order=False))
    est = saltelli.saltelli_indices(Y, N, d, second_order=False)
    s1, _ = ofun.g_function_indices(d)
    st = ofun.g_function_total_exact(d)   # the reference's s_total is off by (1 + V_i), see oracle/functions.py
    np.testing.assert_allclose(est['S1'][0], s1, atol=0.01)
    np.testing.assert_allclose(est['ST'][0], st, atol=0.01)


def test_saltelli_multi_output_aggregation_and_shift_invariance():
    N, p = 5000, 3
    A, B = philox.base_samples(5, 0, N, [('uniform', -np.pi, np.pi)] * p)
    X = saltelli.design_from_base(A, B)
    Y = np.hstack([ofun.ishigami(X), 3.0 * ofun.ishigami(X, a=5.0) + 100.0, X[:, :1] + 0.1 * X[:, 1:2]])
    est = saltelli.saltelli_indices(Y, N, p)
    np.testing.assert_allclose(est['S1_agg'], est['V1'].sum(0) / est['var'].sum())
    shifted = saltelli.saltelli_indices(Y + 1e3, N, p)       # centring makes the estimator shift-invariant
    np.testing.assert_allclose(shifted['S1'], est['S1'], atol=1e-9)
    np.testing.assert_allclose(shifted['ST'], est['ST'], atol=1e-9)
    s2, s1, st = saltelli.uq_sobol_aggregate(Y, N, p)
    assert s2.shape == (p, p) and s1.shape == (p,) and st.shape == (p,)
    blk = saltelli.uq_sobol_aggregate(saltelli.int_func(Y), N, p, indices='block')
    assert blk[1].shape == (p,)


def test_int_func_is_numpy_trapz():
    Y = np.random.RandomState(0).rand(7, 5)
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    np.testing.assert_allclose(saltelli.int_func(Y)[:, 0], [trapz(r) for r in Y], rtol=1e-15)
    assert np.all(saltelli.int_func(Y[:, :1]) == 0.0)      # scalar output: empty sum (reference behaviour)


def test_jackknife_sigma_tracks_the_sampling_spread():
    """The delete-a-group jackknife standard error of the aggregated indices (what the drop-in reports where the
    reference reports OpenTURNS' asymptotic intervals) against the empirical spread over independent replications
    of the experiment (Ishigami, analytic function, N = 2000, 25 groups)."""
    from oracle import functions as ofun
    from oracle import philox, saltelli
    N, p, reps = 2000, 3, 40
    dists = [('uniform', -np.pi, np.pi)] * p
    groups = np.arange(N) % 25
    thetas, sigmas = [], []
    for r in range(reps):
        A, B = philox.base_samples(1000 + r, 0, N, dists)
        Y = ofun.ishigami(saltelli.design_from_base(A, B))
        if r < 6:
            th, sg = saltelli.jackknife_aggregated(Y, N, p, groups)
            sigmas.append(sg)
        else:
            est = saltelli.saltelli_indices(Y, N, p)
            th = np.concatenate([est['S1_agg'], est['ST_agg']])
        thetas.append(th)
    spread = np.std(np.array(thetas), axis=0, ddof=1)
    jk = np.mean(np.array(sigmas), axis=0)
    assert np.all(jk > 0.55 * spread) and np.all(jk < 1.8 * spread), (jk, spread)
