"""GPU parity: batched LML / blocked Cholesky / final factorisation (K5) vs scikit-learn.

Fit parity is defined on LML(theta) at a given theta and on the fitted state at a given theta;
the optimiser path itself is not reproducible in the reference (unseeded DE, kriging.py:170-171).
"""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern, WhiteKernel

from oracle import functions as ofun
from oracle import gp as ogp

pytestmark = pytest.mark.gpu


def _data(n, p, seed=0):
    X = ofun.halton(n, p)
    y = ofun.g_function(X)[:, 0] + 0.01 * np.random.RandomState(seed).randn(n)
    return X, y


@pytest.mark.parametrize('n,p', [(60, 3), (128, 2), (300, 5), (1024, 8)])
def test_lml_batched_matches_sklearn(n, p):
    import torch
    from batman_b200 import kernel_spec
    from batman_b200.fit import normalise_y
    from batman_b200.lml import lml_batched
    X, y = _data(n, p)
    yn, _, _ = normalise_y(y)
    kern = ConstantKernel(1.0) * Matern(length_scale=[1.0] * p, nu=1.5)
    gp = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, y)
    rng = np.random.RandomState(1)
    thetas = [kern.theta] + [kern.theta + rng.uniform(-1.5, 1.5, size=p + 1) for _ in range(6)]
    want = np.array([gp.log_marginal_likelihood(t) for t in thetas])
    specs = [kernel_spec.flatten(kern.clone_with_theta(t), p) for t in thetas]
    got = lml_batched(torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda(), specs)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-7)
    # oracle restatement agrees too
    mine = np.array([ogp.lml(ogp.from_sklearn(kern.clone_with_theta(t)), X, yn) for t in thetas])
    np.testing.assert_allclose(got, mine, rtol=1e-9, atol=1e-7)


def test_lml_other_kernels_and_noise():
    import torch
    from batman_b200 import kernel_spec
    from batman_b200.fit import normalise_y
    from batman_b200.lml import lml_batched
    X, y = _data(200, 4)
    yn, _, _ = normalise_y(y)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda()
    for kern in [ConstantKernel(2.0) * RBF(length_scale=[0.5, 1.0, 1.5, 2.0]) + WhiteKernel(1e-3),
                 Matern(length_scale=0.8, nu=2.5) + WhiteKernel(0.5),
                 ConstantKernel(0.7) + Matern(length_scale=0.5, nu=0.5) + WhiteKernel(1.0)]:
        gp = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, y)
        got = lml_batched(Xd, yd, [kernel_spec.flatten(kern, 4)])
        assert got[0] == pytest.approx(gp.log_marginal_likelihood(kern.theta), rel=1e-9, abs=1e-7)


def test_lml_not_positive_definite_is_minus_inf():
    """sklearn:_gpr.py:591-593: a failed Cholesky gives -inf, and must not poison its batch neighbours."""
    import torch
    from batman_b200 import kernel_spec
    from batman_b200.fit import normalise_y
    from batman_b200.lml import lml_batched
    X, y = _data(150, 3)
    yn, _, _ = normalise_y(y)
    good = kernel_spec.flatten(ConstantKernel(1.0) * Matern(length_scale=[1.0] * 3, nu=1.5), 3)
    bad = kernel_spec.KernelSpec(good.kind, 0.0, -1.0, 1.0, good.ls)       # negative definite off-diagonal mass
    bad2 = kernel_spec.KernelSpec(good.kind, 5.0, 1.0, 1.0, good.ls)       # kdiag << row sums: indefinite
    got = lml_batched(torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda(), [good, bad, good, bad2])
    assert np.isfinite(got[0]) and got[0] == got[2]
    assert got[1] == -np.inf and got[3] == -np.inf


@pytest.mark.parametrize('n,p', [(100, 3), (517, 6)])
def test_factorise_matches_sklearn_state(n, p):
    import torch
    from scipy.linalg import cholesky
    from batman_b200 import Kriging
    X, y = _data(n, p)
    kern = ConstantKernel(1.7, constant_value_bounds='fixed') * Matern(
        length_scale=list(np.linspace(0.6, 1.8, p)), length_scale_bounds='fixed', nu=1.5)
    k = Kriging(X, y[:, None], kernel=kern)             # no free hyper-parameter: factorisation only
    ref = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, y)
    gp = k.gp_models[0]
    np.testing.assert_allclose(gp.L_, ref.L_, rtol=0, atol=1e-11)
    np.testing.assert_allclose(gp.alpha_, ref.alpha_, rtol=1e-6, atol=1e-9 * np.abs(ref.alpha_).max())
    assert gp.log_marginal_likelihood_value_ == pytest.approx(ref.log_marginal_likelihood_value_, rel=1e-10, abs=1e-8)
    pts = np.random.RandomState(3).rand(500, p)
    mean, sigma = k.evaluate(pts)
    rm, rs = ref.predict(pts, return_std=True)
    np.testing.assert_allclose(mean[:, 0], rm, rtol=0, atol=1e-8 * np.abs(rm).max())
    np.testing.assert_allclose(sigma[:, 0] ** 2, rs ** 2, rtol=0, atol=1e-9 * ref._y_train_std ** 2 * 1.7)


def test_global_optimizer_finds_at_least_sklearns_optimum():
    """Default Kriging(): 3 DE runs (vectorised, one batched LML per generation) + polish."""
    from batman_b200 import Kriging
    np.random.seed(123456)
    X = ofun.halton(40, 2, [[0.0, 0.0], [1.0, 1.0]])
    y = ofun.michalewicz(np.pi * X)
    k = Kriging(X, y)
    ref = GaussianProcessRegressor(kernel=ConstantKernel() * Matern(length_scale=(1.0, 1.0),
                                                                    length_scale_bounds=[(0.01, 100)] * 2),
                                   normalize_y=True, n_restarts_optimizer=20, random_state=0).fit(X, y[:, 0])
    assert k.gp_models[0].log_marginal_likelihood_value_ >= ref.log_marginal_likelihood_value_ - 1e-3
    assert k.hyperparameter[0].shape == (3,)
    # LML at the theta we report must be what sklearn computes at that theta
    assert k.gp_models[0].log_marginal_likelihood_value_ == pytest.approx(
        ref.log_marginal_likelihood(k.gp_models[0].kernel_.theta), rel=1e-8, abs=1e-7)


def test_c_abi_lml_and_sobol_stay_inside_their_buffers():
    """bk_lml_batched / bk_gp_factorise with a canary behind the declared workspace (ragged n: 3 block columns, the
    split-K partials, the side-stream kernels), bk_sobol_accumulate with a canary behind the partial slots."""
    import ctypes
    import torch
    from batman_b200 import _lib, kernel_spec
    from batman_b200.fit import normalise_y
    lib = _lib.load()
    n, p, R = 300, 4, 5
    X, y = _data(n, p)
    yn, _, _ = normalise_y(y)
    # the C ABI takes dense row-major buffers (scipy's Halton points are Fortran-ordered)
    y2 = np.cos(3.0 * X).sum(axis=1)                      # a second output: candidates of several outputs share a call
    yn2, _, _ = normalise_y(y2)
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    yd = torch.from_numpy(np.ascontiguousarray(np.stack([yn, yn2]))).cuda()
    targets = [0, 1, 0, 1, 1]
    kern = ConstantKernel(1.3) * Matern(length_scale=[0.8] * p, nu=1.5)
    rng = np.random.RandomState(0)
    specs = [kernel_spec.flatten(kern.clone_with_theta(kern.theta + rng.uniform(-0.5, 0.5, p + 1)), p) for _ in range(R)]
    CAN, pad = -3.5, 1 << 16
    need = lib.bk_lml_workspace(n, R, 0)
    ws = torch.full((need // 8 + 1 + pad,), CAN, dtype=torch.float64, device='cuda')
    lml = torch.full((R + 64,), CAN, dtype=torch.float64, device='cuda')
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.bk_lml_batched(vp(Xd), vp(yd), 2, _lib.as_int_array(targets), n, p, R,
                                  _lib.as_int_array([s.kind for s in specs]),
                                  _lib.as_double_array([s.c0 for s in specs]), _lib.as_double_array([s.c1 for s in specs]),
                                  _lib.as_double_array([s.kdiag for s in specs]),
                                  _lib.as_double_array(np.concatenate([s.ls for s in specs])), vp(lml), vp(ws), need, None))
    torch.cuda.synchronize()
    assert bool((ws[need // 8 + 1:] == CAN).all()) and bool((lml[R:] == CAN).all())
    gps = [GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, yy) for yy in (y, y2)]
    want = [gps[t].log_marginal_likelihood(np.log(np.r_[s.c1, s.ls])) for s, t in zip(specs, targets)]
    np.testing.assert_allclose(lml[:R].cpu().numpy(), want, rtol=1e-9, atol=1e-7)


def test_c_abi_sobol_partials_canary():
    import ctypes
    import torch
    from conftest import kriging_from_states, load_golden
    from batman_b200 import _lib
    lib = _lib.load()
    z, states = load_golden('kernel_default_3out')
    model = kriging_from_states(states)._model()
    p, m = 3, 3
    nsums, nslots = lib.bk_sobol_nsums(p, 1), lib.bk_sobol_nslots()
    CAN, pad = -1.75, 4096
    part = torch.full((nslots * m * nsums * 2 + pad,), CAN, dtype=torch.float64, device='cuda')
    part[:nslots * m * nsums * 2] = 0.0
    kinds, params = _lib.as_int_array([0] * p), _lib.as_double_array([0.0, 1.0, 0.0, 0.0] * p)
    _lib.check(lib.bk_sobol_accumulate(model.handle, None, None, 123456, kinds, params, 0, 12345, 1,
                                       _lib.as_double_array(model.y_means), ctypes.c_void_p(part.data_ptr()), None))
    torch.cuda.synchronize()
    assert bool((part[nslots * m * nsums * 2:] == CAN).all())
    assert bool(torch.isfinite(part[:nslots * m * nsums * 2]).all())
