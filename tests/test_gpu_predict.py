"""GPU parity: Kriging.evaluate / SurrogateModel.__call__ through the C ABI vs the oracle.

Tolerances (north_star): predictive mean and variance within 1e-10, measured
relative to the output scale -- |dmean| <= 1e-10 * max(|mean|, y_std) and
|dvar| <= 1e-10 * prior variance (kdiag * y_std^2) -- because near training
points sigma -> 0 and sqrt amplifies (SURVEY.md §7.2 "Parity floor").
"""
import numpy as np
import pytest

from conftest import GP_CASES, kriging_from_states, load_golden, synthetic_state
from oracle import gp as ogp
from oracle import gp_hp

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _check(mean, sigma, ref_mean, ref_sigma, states, rtol=RTOL):
    assert mean.shape == ref_mean.shape and sigma.shape == ref_sigma.shape
    for q, st in enumerate(states):
        scale = np.maximum(np.abs(ref_mean[:, q]), st['y_std'])
        assert np.max(np.abs(mean[:, q] - ref_mean[:, q]) / scale) <= rtol
        prior = ogp.kernel_diag(st['tree']) * st['y_std'] ** 2
        assert np.max(np.abs(sigma[:, q] ** 2 - ref_sigma[:, q] ** 2)) <= rtol * prior


@pytest.mark.parametrize('case', GP_CASES)
def test_golden_parity_with_reference_evaluate(case):
    """Device evaluate == batman.surrogate.Kriging.evaluate on identical fitted state."""
    z, states = load_golden(case)
    k = kriging_from_states(states)
    mean, sigma = k.evaluate(z['points'])
    _check(mean, sigma, z['ref_mean'], z['ref_sigma'], states)
    omean, osigma = ogp.evaluate(states, z['points'])
    _check(mean, sigma, omean, osigma, states)


def test_docstring_known_answer():
    """kriging.py:14-19."""
    z, states = load_golden('kriging_docstring')
    mean, sigma = kriging_from_states(states).evaluate((5.0, 8.0))
    assert mean.shape == (1, 2) and sigma.shape == (1, 2)          # multi_eval contract for one 1-D point
    np.testing.assert_allclose(mean[0], [10.333, 3.591], atol=5e-4)
    np.testing.assert_allclose(sigma[0], [1.247, 0.694], atol=5e-4)


def test_ishigami_surrogate_model_call():
    """SurrogateModel.__call__ with the MinMax scaler fused into the kernel (cfg1 model)."""
    from batman_b200 import SurrogateModel
    z, states = load_golden('ishigami_n100')
    s = SurrogateModel('kriging', z['corners'], ['x1', 'x2', 'x3'])
    s.adopt(kriging_from_states(states), z['X_raw'])
    mean, sigma = s(z['points'])
    _check(mean, sigma, z['ref_mean'], z['ref_sigma'], states)
    m1, s1 = s(z['points'][3])                                     # single point
    assert m1.shape == (1, 1)
    np.testing.assert_allclose(m1[0], mean[3], rtol=0, atol=1e-13)


def test_channel_flow_pod_surrogate():
    """POD inverse transform applied to mean AND sigma (surrogate_model.py:191-196)."""
    from batman_b200 import Pod, SurrogateModel
    z, states = load_golden('channel_flow_pod')
    s = SurrogateModel('kriging', z['corners'], ['Ks', 'Q'])
    s.adopt(kriging_from_states(states), z['X_raw'], pod=Pod(z['mean_snapshot'], z['U']))
    mean, sigma = s(z['points'])
    assert mean.shape == z['ref_mean'].shape
    np.testing.assert_allclose(mean, z['ref_mean'], rtol=0, atol=1e-10 * np.abs(z['ref_mean']).max())
    np.testing.assert_allclose(sigma, z['ref_sigma'], rtol=0, atol=1e-7 * np.abs(z['ref_sigma']).max())


EPS = np.finfo(np.float64).eps


@pytest.mark.parametrize('n,p,kind,ls', [(1024, 8, 'matern32', 0.7), (1024, 8, 'matern32', 3.0),
                                         (700, 5, 'matern52', 1.0), (333, 3, 'rbf', 0.12),
                                         (256, 10, 'matern12', 2.0), (513, 20, 'rbf', 2.5),
                                         (333, 3, 'rbf', 0.4), (4096, 10, 'matern32', 1.0)])
def test_synthetic_parity_vs_oracle_and_longdouble(n, p, kind, ls):
    """Sizes the oracle finishes in seconds (n = 4096 is BASELINE config 3's model size); ragged n (not a
    multiple of 128) and k (not of 256).

    The mean is a sign-alternating sum sum_j k_j alpha_j; its conditioning
    A = sum|k_j alpha_j| * y_std / scale is reported next to the error.  The
    gate is 1e-10 relative to the output scale, widened to the conditioning floor
    8 eps A only where A itself exceeds 1e-10 / 8 eps (numerically singular K, e.g. the
    last RBF case with cond(K) ~ 1/jitter, where sklearn itself is that far from the
    long-double value)."""
    st = synthetic_state(n, p, kind, ls=np.full(p, ls))
    k = kriging_from_states([st])
    rng = np.random.RandomState(42)
    pts = np.vstack([rng.rand(1000, p), st['X_train'][:5], st['X_train'][:5] + 1e-7])
    mean, sigma = k.evaluate(pts)
    omean, osigma = ogp.evaluate([st], pts)
    scale = np.maximum(np.abs(omean[:, 0]), st['y_std'])
    amp = gp_hp.amplification(st, pts) * np.abs(omean[:, 0] - st['y_mean']) / scale
    gate = np.maximum(RTOL, 8 * EPS * amp)
    err = np.abs(mean[:, 0] - omean[:, 0]) / scale
    assert np.all(err <= gate), (err.max(), amp.max())
    prior = ogp.kernel_diag(st['tree']) * st['y_std'] ** 2
    assert np.max(np.abs(sigma[:, 0] ** 2 - osigma[:, 0] ** 2)) <= RTOL * prior
    # arbiter: the long-double evaluation of the same formulas on the same float64 state
    sub = slice(0, 48)
    mld, sld = gp_hp.predict_hp(st, pts[sub])
    err_gpu = np.abs(mean[sub, 0] - mld.astype(float)) / scale[sub]
    err_sk = np.abs(omean[sub, 0] - mld.astype(float)) / scale[sub]
    print("n=%d p=%d %s ls=%g: max amp %.2e, GPU-vs-longdouble %.2e, sklearn-vs-longdouble %.2e"
          % (n, p, kind, ls, amp.max(), err_gpu.max(), err_sk.max()))
    assert np.all(err_gpu <= gate[sub]), (err_gpu.max(), err_sk.max(), amp.max())
    assert np.max(np.abs(sigma[sub, 0] ** 2 - (sld ** 2).astype(float))) <= RTOL * prior


def test_interpolation_property_at_training_points():
    """Size-independent property: a noise-free GP reproduces its data, sigma ~ 0 there."""
    st = synthetic_state(512, 4, 'matern32', ls=np.full(4, 0.5))
    from oracle import functions as ofun
    y = ofun.g_function(st['X_train'])[:, 0]
    mean, sigma = kriging_from_states([st]).evaluate(st['X_train'])
    assert np.max(np.abs(mean[:, 0] - y)) < 1e-6
    assert np.max(sigma) < 1e-3 * st['y_std']


def test_edge_cases():
    z, states = load_golden('kernel_default_3out')
    k = kriging_from_states(states)
    with pytest.raises(ValueError):
        k.evaluate(np.zeros((4, 7)))                               # wrong feature count
    mean = k.evaluate(z['points'][:7], return_std=False)           # mean only path
    m2, _ = k.evaluate(z['points'][:7])
    np.testing.assert_array_equal(mean, m2)
    big = np.tile(z['points'], (40, 1))                            # several CTAs, K* chunking
    mb, sb = k.evaluate(big)
    m0, s0 = k.evaluate(z['points'])
    np.testing.assert_array_equal(mb[:len(m0)], m0)
    np.testing.assert_array_equal(sb[-len(s0):], s0)


def test_pipelined_host_path_equals_single_shot(monkeypatch):
    """Large numpy batches go through the copy/compute pipeline: results must be bit-identical."""
    from batman_b200.device_model import DeviceModel
    z, states = load_golden('kernel_default_3out')
    k = kriging_from_states(states)
    pts = np.tile(z['points'], (60, 1))
    m0, s0 = k.evaluate(pts)
    monkeypatch.setattr(DeviceModel, 'E2E_CHUNK', 700)          # 3360 rows -> 5 slices, last one ragged
    m1, s1 = k.evaluate(pts)
    np.testing.assert_array_equal(m0, m1)
    np.testing.assert_array_equal(s0, s1)
    m2 = k.evaluate(pts, return_std=False)
    np.testing.assert_array_equal(m0, m2)


def test_pickle_roundtrip_drops_device_buffers():
    import pickle
    z, states = load_golden('kernel_default_3out')
    k = kriging_from_states(states)
    m0, s0 = k.evaluate(z['points'])
    k2 = pickle.loads(pickle.dumps(k))
    assert k2._device is None
    m1, s1 = k2.evaluate(z['points'])
    np.testing.assert_array_equal(m0, m1)
    np.testing.assert_array_equal(s0, s1)


def _child_evaluate(blob, pts, q):
    import pickle
    k = pickle.loads(blob)
    q.put(k.evaluate(pts))


def test_pickled_model_evaluates_in_a_fresh_process():
    """SURVEY §8(b) threading note: the reference ships fitted predictors to pool workers (pod.py:366-411,
    surrogate_model.py:236-258).  CUDA contexts do not survive fork, so device state is created lazily in the
    process that evaluates; a spawned worker unpickles the model and gets bit-identical results."""
    import multiprocessing as mp
    import pickle
    z, states = load_golden('kernel_default_3out')
    k = kriging_from_states(states)
    m0, s0 = k.evaluate(z['points'])
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    proc = ctx.Process(target=_child_evaluate, args=(pickle.dumps(k), z['points'], q))
    proc.start()
    m1, s1 = q.get(timeout=300)
    proc.join(60)
    assert proc.exitcode == 0
    np.testing.assert_array_equal(m0, m1)
    np.testing.assert_array_equal(s0, s1)


def test_fit_on_device_reproduces_docstring_and_sklearn_lml():
    """Kriging(...) fits on the device: LML(theta*) must agree with sklearn's LML at the same theta."""
    from batman_b200 import Kriging
    from sklearn.gaussian_process import GaussianProcessRegressor
    np.random.seed(123456)
    sample = np.array([[2, 4], [3, 5], [6, 9]], dtype=float)
    data = np.array([[12, 1], [10, 2], [9, 4]], dtype=float)
    k = Kriging(sample, data, global_optimizer=False)
    mean, sigma = k.evaluate((5.0, 8.0))
    np.testing.assert_allclose(mean[0], [10.333, 3.591], atol=2e-2)
    np.testing.assert_allclose(sigma[0], [1.247, 0.694], atol=2e-2)
    for q, gp in enumerate(k.gp_models):
        ref = GaussianProcessRegressor(kernel=gp.kernel_, normalize_y=True, optimizer=None).fit(sample, data[:, q])
        assert gp.log_marginal_likelihood_value_ == pytest.approx(ref.log_marginal_likelihood_value_, rel=1e-10, abs=1e-10)
        np.testing.assert_allclose(gp.alpha_, ref.alpha_, rtol=1e-8)
        rm, rs = ref.predict(np.array([[5.0, 8.0]]), return_std=True)
        assert mean[0, q] == pytest.approx(rm[0], rel=1e-9)
        assert sigma[0, q] == pytest.approx(rs[0], rel=1e-7)


@pytest.mark.parametrize('scheme', [0, 1])
@pytest.mark.parametrize('n,p,kind,ls', [(1024, 8, 'matern32', 3.0), (700, 5, 'matern52', 1.0), (100, 3, 'matern32', 0.7),
                                         (333, 3, 'rbf', 0.4)])
def test_int8_tensor_core_sigma_matches_fp64_path(n, p, kind, ls, scheme):
    """k_var_ozaki (tcgen05 int8, error-free slicing; scheme 0 = 8 x 7-bit, 1 = 7 x 8-bit with unsigned K* bytes)
    against k_var_trmm (FP64 DMMA) and the long-double value."""
    st = synthetic_state(n, p, kind, ls=np.full(p, ls))
    k = kriging_from_states([st])
    rng = np.random.RandomState(5)
    pts = np.vstack([rng.rand(700, p), st['X_train'][:5]])
    model = k._model()
    model.set_sigma_kernel('dmma')
    m0, s0 = k.evaluate(pts)
    model.set_sigma_kernel('int8')
    model.set_int8_scheme(scheme)
    m1, s1 = k.evaluate(pts)
    prior = ogp.kernel_diag(st['tree']) * st['y_std'] ** 2
    np.testing.assert_array_equal(m0, m1)
    err = np.max(np.abs(s1 ** 2 - s0 ** 2)) / prior
    _, sld = gp_hp.predict_hp(st, pts[:32])
    err_ld = np.max(np.abs(s1[:32, 0] ** 2 - (sld ** 2).astype(float))) / prior
    print("n=%d %s scheme %d: int8-vs-DMMA %.2e, int8-vs-longdouble %.2e (of the prior variance)" % (n, kind, scheme, err, err_ld))
    assert err <= 1e-11 and err_ld <= RTOL


def test_int8_slicing_at_the_top_of_the_covariance_range():
    """k(x, x) = c reaches the bound the K* slicing is scaled by; with c one ulp below a power of two the scaled
    value is 1 - 2^-53: the slices must not wrap.  Sigma at and next to training points, int8 vs FP64 kernel."""
    c = np.nextafter(2.0, 0.0)
    st = synthetic_state(300, 3, 'matern32', ls=np.full(3, 0.6), c=c)
    k = kriging_from_states([st])
    pts = np.vstack([st['X_train'][:64], st['X_train'][:64] + 1e-9, np.random.RandomState(0).rand(200, 3)])
    model = k._model()
    model.set_sigma_kernel('dmma')
    m0, s0 = k.evaluate(pts)
    for scheme in (0, 1):
        model.set_int8_scheme(scheme)
        model.set_sigma_kernel('int8')
        m1, s1 = k.evaluate(pts)
        assert np.array_equal(m0, m1)
        assert np.max(np.abs(s1 ** 2 - s0 ** 2)) <= 1e-12 * c * st['y_std'] ** 2


def test_c_abi_predict_stays_inside_its_buffers():
    """bk_gp_predict called directly with canary-framed output and workspace buffers (ragged k, both sigma kernels,
    a fast-path and a general-tree output): nothing outside [k x m] results and the declared workspace is touched."""
    import ctypes
    import torch
    from batman_b200 import _lib
    from oracle import functions as ofun
    lib = _lib.load()
    p, n = 3, 200
    X = ofun.halton(n, p)
    y0 = ofun.g_function(X)[:, 0]
    y1 = np.sin(3 * X[:, 0]) + X[:, 1] * X[:, 2]
    ls = np.full(p, 0.7)
    states = [ogp.fit_state(('prod', ('const', 1.5), ('matern', 1.5, ls)), X, y0),
              ogp.fit_state(('sum', ('prod', ('const', 1.2), ('rbf', ls)), ('prod', ('const', 0.4), ('matern', 0.5, 2 * ls))), X, y1)]
    kr = kriging_from_states(states)
    model = kr._model()
    model.ensure_w()
    k, m, pad = 777, 2, 4096
    pts = torch.rand((k, p), dtype=torch.float64, device='cuda')
    CAN = -7.25
    for kernel in ('dmma', 'int8'):
        model.set_sigma_kernel(kernel)
        need = lib.bk_gp_predict_workspace(model.handle, k, 1)
        mean = torch.full((pad + k * m + pad,), CAN, dtype=torch.float64, device='cuda')
        std = torch.full((pad + k * m + pad,), CAN, dtype=torch.float64, device='cuda')
        ws = torch.full((need // 8 + 1 + pad,), CAN, dtype=torch.float64, device='cuda')
        ws[:need // 8 + 1] = 0.0
        _lib.check(lib.bk_gp_predict(model.handle, ctypes.c_void_p(pts.data_ptr()), k,
                                     ctypes.c_void_p(mean.data_ptr() + pad * 8), ctypes.c_void_p(std.data_ptr() + pad * 8),
                                     ctypes.c_void_p(ws.data_ptr()), need, None))
        torch.cuda.synchronize()
        for buf in (mean, std):
            assert bool((buf[:pad] == CAN).all()) and bool((buf[pad + k * m:] == CAN).all())
            assert not bool((buf[pad:pad + k * m] == CAN).any())
        assert bool((ws[need // 8 + 1:] == CAN).all())
        got_m = mean[pad:pad + k * m].reshape(k, m).cpu().numpy()
        got_s = std[pad:pad + k * m].reshape(k, m).cpu().numpy()
        om, osd = ogp.evaluate(states, pts.cpu().numpy())
        _check(got_m, got_s, om, osd, states)


def test_overlap_lanes_give_identical_results():
    """bk_gp_set_overlap(1): K1 of unit u+1 beside the sigma kernel of unit u on two internal streams and two K*
    buffers -- same bits as the single-stream path, for several chunks (small workspace) x two outputs, both sigma
    kernels, called twice in a row on the caller's stream (the fork/join must order consecutive calls)."""
    import ctypes
    import torch
    from batman_b200 import _lib
    from oracle import functions as ofun
    lib = _lib.load()
    p, n = 4, 300
    X = ofun.halton(n, p)
    ls = np.full(p, 0.6)
    states = [ogp.fit_state(('prod', ('const', 1.3), ('matern', 1.5, ls)), X, ofun.g_function(X)[:, 0]),
              ogp.fit_state(('prod', ('const', 0.8), ('matern', 2.5, 1.5 * ls)), X, np.sin(4 * X[:, 0]) + X[:, 1])]
    model = kriging_from_states(states)._model()
    model.ensure_w()
    k = 3000
    pts = torch.rand((k, p), dtype=torch.float64, device='cuda')
    per_point = model.n_pad * 8
    ws_bytes = 2 * 1024 * per_point + 256                   # two K* buffers of 1024 points: 3 chunks x 2 outputs = 6 units
    ws = torch.zeros(ws_bytes // 8 + 1, dtype=torch.float64, device='cuda')
    for kernel in ('int8', 'dmma'):
        model.set_sigma_kernel(kernel)
        res = []
        for on in (0, 1, 1):
            _lib.check(lib.bk_gp_set_overlap(model.handle, on))
            mean = torch.zeros((k, 2), dtype=torch.float64, device='cuda')
            std = torch.zeros((k, 2), dtype=torch.float64, device='cuda')
            _lib.check(lib.bk_gp_predict(model.handle, ctypes.c_void_p(pts.data_ptr()), k, ctypes.c_void_p(mean.data_ptr()),
                                         ctypes.c_void_p(std.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws_bytes, None))
            fault = ctypes.c_int(-1)
            _lib.check(lib.bk_gp_predict_status(ctypes.c_void_p(ws.data_ptr()), ws_bytes, ctypes.byref(fault), None))
            assert fault.value == 0
            res.append((mean.cpu().numpy(), std.cpu().numpy()))
        _lib.check(lib.bk_gp_set_overlap(model.handle, 0))
        for m_on, s_on in res[1:]:
            assert np.array_equal(res[0][0], m_on) and np.array_equal(res[0][1], s_on)
        om, osd = ogp.evaluate(states, pts.cpu().numpy())
        _check(res[1][0], res[1][1], om, osd, states)


def test_pipelined_host_path_matches_single_call():
    """DeviceModel.predict above E2E_CHUNK rows (slices uploaded, predicted and downloaded on three streams): same
    bits as one resident call, ragged last slice, with and without sigma."""
    import torch
    from batman_b200.device_model import DeviceModel
    from oracle import functions as ofun
    p, n = 3, 128
    X = ofun.halton(n, p)
    states = [ogp.fit_state(('prod', ('const', 1.1), ('matern', 1.5, np.full(p, 0.5))), X, ofun.g_function(X)[:, 0])]
    model = kriging_from_states(states)._model()
    k = 2 * DeviceModel.E2E_CHUNK + 12345
    rng = np.random.RandomState(3)
    pts = rng.rand(k, p)
    mean, std = model.predict(pts, return_std=True)
    dm, ds = model.predict_device(torch.from_numpy(pts).cuda(), return_std=True)
    assert np.array_equal(mean, dm.cpu().numpy()) and np.array_equal(std, ds.cpu().numpy())
    mean2, none = model.predict(pts, return_std=False)
    assert none is None and np.array_equal(mean2, mean)


@pytest.mark.parametrize('scheme', [0, 1])
def test_int8_slicing_of_w_rows_with_wide_dynamic_range(scheme):
    """The int8 path slices every row of W = L^-1 in fixed point relative to the row's largest entry.  A
    near-singular K (short-range RBF, pairs of training points 1e-5 apart, cond(K) ~ 1/jitter) gives rows of W whose entries span
    more than 2^40; the variance from the slices must then be as close to the long-double value as the FP64 DMMA
    kernel's is (56 bits of fixed point against 53 bits of floating point per product, exact accumulation against
    rounded accumulation)."""
    import torch
    from oracle import functions as ofun
    rng = np.random.RandomState(5)
    p, n0 = 2, 160
    X0 = ofun.halton(n0, p)
    X = np.vstack([X0, X0[:48] + 1e-5 * rng.randn(48, p)])
    y = np.sin(5 * X[:, 0]) * np.cos(3 * X[:, 1])
    st = ogp.fit_state(('prod', ('const', 1.0), ('rbf', np.full(p, 0.07))), X, y)
    rows = np.abs(np.linalg.inv(st['L']))
    span = [rows[i, :i + 1].max() / rows[i, :i + 1][rows[i, :i + 1] > 0].min() for i in range(1, len(rows))]
    assert max(span) > 2.0 ** 40, "the case must exercise a wide row"
    k = kriging_from_states([st])
    model = k._model()
    pts = np.vstack([rng.rand(96, p), X[:16] + 3e-6, X[n0:n0 + 16] - 2e-6])
    model.set_sigma_kernel('dmma')
    _, s_f = k.evaluate(pts)
    model.set_int8_scheme(scheme)
    model.set_sigma_kernel('int8')
    _, s_i = k.evaluate(pts)
    _, sld = gp_hp.predict_hp(st, pts)
    ref = (sld ** 2).astype(float)
    prior = ogp.kernel_diag(st['tree']) * st['y_std'] ** 2
    err_i = np.max(np.abs(s_i[:, 0] ** 2 - ref)) / prior
    err_f = np.max(np.abs(s_f[:, 0] ** 2 - ref)) / prior
    print("max |W| %.2e, widest row span 2^%.0f: int8 scheme %d vs long double %.2e, FP64 DMMA vs long double %.2e "
          "(of the prior variance)" % (rows.max(), np.log2(max(span)), scheme, err_i, err_f))
    assert err_i <= max(4.0 * err_f, 1e-12)
    torch.cuda.synchronize()


@pytest.mark.parametrize('n,p,gang', [(2000, 6, 4), (8100, 6, 16)])
def test_int8_sigma_gang_sizes_between_the_configs(n, p, gang):
    """The sigma kernel shares a tile of test points between G CTAs, G chosen from n_train (2 at n = 1024, 8 at 4096,
    37 at 16384: covered by the config tests).  The sizes in between (G = 4, 16), several tiles per gang and a
    ragged last tile: int8 kernel against the FP64 DMMA kernel."""
    st = synthetic_state(n, p, 'matern32', ls=np.full(p, 1.2))
    k = kriging_from_states([st])
    rng = np.random.RandomState(11)
    pts = np.vstack([rng.rand(9001, p), st['X_train'][:7]])
    model = k._model()
    model.set_sigma_kernel('dmma')
    m0, s0 = k.evaluate(pts)
    model.set_sigma_kernel('int8')
    m1, s1 = k.evaluate(pts)
    prior = ogp.kernel_diag(st['tree']) * st['y_std'] ** 2
    np.testing.assert_array_equal(m0, m1)
    err = np.max(np.abs(s1 ** 2 - s0 ** 2)) / prior
    print("n=%d (gang of %d): int8-vs-DMMA %.2e of the prior variance" % (n, gang, err))
    assert err <= 1e-11
