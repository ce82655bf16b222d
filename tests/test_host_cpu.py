"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
from sklearn.gaussian_process import kernels as sk

from conftest import ROOT
from batman_b200 import _lib, distributions, kernel_spec
from batman_b200.surrogate_model import MinMax


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'batman_b200.h')).read()
    declared = set(re.findall(r'\b(bk_[a-z_0-9]+)\s*\(', header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libbatman_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_library_host_only_queries():
    lib = _lib.load()
    assert lib.bk_version() >= 100
    assert lib.bk_gp_padded_dim(3) == 4 and lib.bk_gp_padded_dim(8) == 8 and lib.bk_gp_padded_dim(33) == -1
    assert lib.bk_gp_padded_n(100) == 128 and lib.bk_gp_padded_n(1024) == 1024
    assert lib.bk_sobol_nsums(3, 1) == 5 + 18 + 6 + 3 + 1      # + 1: the row count of the slot
    assert lib.bk_sobol_nsums(2, 1) == 5 + 12 + 1          # p == 2: C blocks are E blocks
    assert lib.bk_sobol_nsums(8, 0) == 5 + 48 + 1
    h = ctypes.c_void_p()
    assert lib.bk_gp_create(ctypes.byref(h), 10, 40, 1) != 0      # p too large: error, not a crash
    assert b'dimension' in lib.bk_last_error()
    assert lib.bk_gp_create(ctypes.byref(h), 10, 3, 2) == 0
    assert lib.bk_gp_set_output(h, 5, 1, 0.0, 1.0, 1.0, None, None, None, None, 0.0, 1.0) != 0
    assert lib.bk_gp_destroy(h) == 0


def test_kernel_flatten_matches_sklearn_semantics():
    rng = np.random.RandomState(0)
    X, Y = rng.rand(7, 3), rng.rand(5, 3)
    cases = [
        sk.ConstantKernel(2.5) * sk.Matern(length_scale=[0.3, 1.0, 2.0], nu=1.5),
        sk.Matern(length_scale=0.7, nu=1.5) + sk.WhiteKernel(0.8),
        sk.ConstantKernel(1.3) + sk.Matern(length_scale=0.5, nu=0.5) + sk.WhiteKernel(1.0),
        sk.ConstantKernel(0.5) * sk.RBF(length_scale=[1.0, 2.0, 3.0]) + sk.WhiteKernel(1e-3),
        sk.ConstantKernel(2.0) * sk.Matern(length_scale=1.1, nu=2.5) * sk.ConstantKernel(3.0),
        sk.Matern(length_scale=1.1, nu=np.inf),
    ]
    stat = {kernel_spec.KIND_MATERN12: lambda r: np.exp(-r),
            kernel_spec.KIND_MATERN32: lambda r: (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r),
            kernel_spec.KIND_MATERN52: lambda r: (1 + np.sqrt(5) * r + 5 * r * r / 3) * np.exp(-np.sqrt(5) * r),
            kernel_spec.KIND_RBF: lambda r: np.exp(-0.5 * r * r),
            kernel_spec.KIND_MATERN_INF: lambda r: np.exp(-0.5 * r * r)}
    for k in cases:
        spec = kernel_spec.flatten(k, 3)
        r = np.sqrt((((X[:, None, :] - Y[None, :, :]) / spec.ls) ** 2).sum(-1))
        np.testing.assert_allclose(spec.c0 + spec.c1 * stat[spec.kind](r), k(X, Y), rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(np.full(7, spec.kdiag), k.diag(X), rtol=1e-15)


def test_kernel_flatten_general_trees():
    """Trees with several stationary leaves expand to sum_t coef_t prod_f stat_f^e_tf (SURVEY §8 A5 note)."""
    rng = np.random.RandomState(1)
    X, Y = rng.rand(7, 3), rng.rand(5, 3)
    cases = [
        sk.ConstantKernel(2.0) * sk.RBF(length_scale=[1.0, 2.0, 0.5]) * sk.Matern(length_scale=0.7, nu=2.5),
        sk.ConstantKernel(1.5) * sk.RBF(length_scale=1.0) + sk.ConstantKernel(0.5) * sk.Matern(length_scale=2.0, nu=0.5)
        + sk.WhiteKernel(0.01),
        (sk.ConstantKernel(2.0) + sk.RBF(length_scale=1.0)) * (sk.ConstantKernel(3.0) + sk.Matern(length_scale=0.4)),
        sk.RBF(length_scale=1.3) * sk.RBF(length_scale=1.3),
        sk.Matern(length_scale=1.0, nu=0.5) + sk.Matern(length_scale=2.0, nu=1.5) + sk.Matern(length_scale=3.0, nu=np.inf),
    ]
    stat = {kernel_spec.KIND_MATERN12: lambda r: np.exp(-r),
            kernel_spec.KIND_MATERN32: lambda r: (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r),
            kernel_spec.KIND_MATERN52: lambda r: (1 + np.sqrt(5) * r + 5 * r * r / 3) * np.exp(-np.sqrt(5) * r),
            kernel_spec.KIND_RBF: lambda r: np.exp(-0.5 * r * r),
            kernel_spec.KIND_MATERN_INF: lambda r: np.exp(-0.5 * r * r)}
    for k in cases:
        spec = kernel_spec.flatten(k, 3)
        g = spec.general
        assert spec.kind == kernel_spec.KIND_GENERAL and g is not None
        fv = [stat[g.fkind[f]](np.sqrt((((X[:, None, :] - Y[None, :, :]) / g.fls[f]) ** 2).sum(-1))) for f in range(g.nf)]
        val = sum(g.tcoef[t] * np.prod([fv[f] ** g.texp[t, f] for f in range(g.nf)], axis=0) for t in range(g.nt))
        np.testing.assert_allclose(val, k(X, Y), rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(np.full(7, spec.kdiag), k.diag(X), rtol=1e-15)
        assert spec.nonneg


def test_kernel_flatten_random_trees_match_sklearn():
    """200 random Sum/Product trees over Constant / White / RBF / Matern leaves: the expanded spec either evaluates
    to sklearn's k(X, Y) and diag(X), or raises NotImplementedError (expansion beyond 3 leaves / 6 terms / power 4)."""
    rng = np.random.RandomState(7)
    X, Y = rng.rand(6, 2), rng.rand(5, 2)
    stat = {kernel_spec.KIND_MATERN12: lambda r: np.exp(-r),
            kernel_spec.KIND_MATERN32: lambda r: (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r),
            kernel_spec.KIND_MATERN52: lambda r: (1 + np.sqrt(5) * r + 5 * r * r / 3) * np.exp(-np.sqrt(5) * r),
            kernel_spec.KIND_RBF: lambda r: np.exp(-0.5 * r * r),
            kernel_spec.KIND_MATERN_INF: lambda r: np.exp(-0.5 * r * r)}
    pool = [lambda: sk.RBF(length_scale=[0.7, 1.3]), lambda: sk.Matern(length_scale=0.9, nu=1.5),
            lambda: sk.Matern(length_scale=[1.1, 0.4], nu=2.5), lambda: sk.Matern(length_scale=2.0, nu=0.5)]

    def tree(depth):
        u = rng.rand()
        if depth == 0 or u < 0.3:
            v = rng.rand()
            if v < 0.25:
                return sk.ConstantKernel(float(rng.uniform(0.2, 3.0)))
            if v < 0.35:
                return sk.WhiteKernel(float(rng.uniform(1e-3, 0.5)))
            return pool[rng.randint(len(pool))]()
        a, b = tree(depth - 1), tree(depth - 1)
        return sk.Sum(a, b) if rng.rand() < 0.5 else sk.Product(a, b)

    n_general = n_affine = n_rejected = 0
    for _ in range(200):
        k = tree(3)
        try:
            spec = kernel_spec.flatten(k, 2)
        except NotImplementedError:
            n_rejected += 1
            continue
        if spec.general is None:
            n_affine += 1
            r = np.sqrt((((X[:, None, :] - Y[None, :, :]) / spec.ls) ** 2).sum(-1))
            val = spec.c0 + spec.c1 * stat[spec.kind](r)
        else:
            n_general += 1
            g = spec.general
            assert g.nf <= 3 and g.nt <= 6 and g.texp.max() <= 4
            fv = [stat[g.fkind[f]](np.sqrt((((X[:, None, :] - Y[None, :, :]) / g.fls[f]) ** 2).sum(-1)))
                  for f in range(g.nf)]
            val = sum(g.tcoef[t] * np.prod([fv[f] ** g.texp[t, f] for f in range(g.nf)], axis=0) for t in range(g.nt))
        np.testing.assert_allclose(val, k(X, Y), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(np.full(6, spec.kdiag), k.diag(X), rtol=1e-14)
    assert n_general > 20 and n_affine > 20 and n_rejected > 5, (n_general, n_affine, n_rejected)


def test_kernel_flatten_rejects_unsupported():
    four = sk.Matern(length_scale=1.0) + sk.RBF(length_scale=2.0) + sk.Matern(length_scale=3.0, nu=0.5) \
        + sk.Matern(length_scale=4.0, nu=2.5)                           # 4 distinct stationary leaves
    for k in [sk.RationalQuadratic(), sk.Matern(nu=1.0), sk.DotProduct(), four,
              sk.ConstantKernel(2.0) * sk.ExpSineSquared()]:
        with pytest.raises(NotImplementedError):
            kernel_spec.flatten(k, 2)
    with pytest.raises(ValueError):
        kernel_spec.flatten(sk.Matern(length_scale=[1.0, 2.0]), 3)


def test_distributions():
    assert distributions.parse('Uniform(-3.1415, 3.1415)') == (0, -3.1415, 3.1415, 0.0, 0.0)
    assert distributions.parse('Normal(4035., 400.)') == (1, 4035.0, 400.0, 0.0, 0.0)
    assert distributions.parse(('uniform', 0, 1)) == (0, 0.0, 1.0, 0.0, 0.0)
    assert distributions.parse('Uniform()') == (0, -1.0, 1.0, 0.0, 0.0)              # OpenTURNS defaults
    with pytest.raises(SystemError):                 # uq.py:156-158
        distributions.parse('Unifrom(0, 1)')
    with pytest.raises(SystemError):                 # a parameter object without .getDistribution() is not a distribution
        distributions.parse('BetaMuSigma(4035, 400, 2500, 6000)')
    with pytest.raises(SystemError):
        distributions.parse('Uniform(3, 1)')
    with pytest.raises(NotImplementedError):         # a real OpenTURNS family without a device inverse CDF
        distributions.parse('Gamma(2., 1., 0.)')
    kinds, params = distributions.parse_all(['Uniform(15., 60.)', 'Normal(4035., 400.)'])
    assert kinds == [0, 1] and params == [15.0, 60.0, 0.0, 0.0, 4035.0, 400.0, 0.0, 0.0]


def test_distribution_parametrisations_against_scipy_stats():
    """The host statement of every inverse CDF equals scipy.stats' ppf of the distribution the OpenTURNS 1.15
    constructor describes (an independent check of parametrisation and formula); the device twins are compared
    with the same formulas in tests/test_gpu_sobol.py."""
    from scipy import stats
    u = np.linspace(0.001, 0.999, 499)
    cases = [
        ('LogNormal(0.3, 0.6, 1.5)', stats.lognorm(s=0.6, loc=1.5, scale=np.exp(0.3))),
        ('Triangular(1., 2., 5.)', stats.triang(c=0.25, loc=1.0, scale=4.0)),
        ('TruncatedNormal(1., 2., -1., 2.5)', stats.truncnorm(a=-1.0, b=0.75, loc=1.0, scale=2.0)),
        ('Beta(2.5, 4., -1., 3.)', stats.beta(2.5, 4.0, loc=-1.0, scale=4.0)),
        ('Gumbel(2., 1.)', stats.gumbel_r(loc=1.0, scale=2.0)),
        ('Exponential(0.5, 2.)', stats.expon(loc=2.0, scale=2.0)),
        ('LogUniform(-1., 2.)', stats.loguniform(np.exp(-1.0), np.exp(2.0))),
        ('WeibullMin(2., 1.5, 0.5)', stats.weibull_min(c=1.5, loc=0.5, scale=2.0)),
        ('Logistic(1., 2.)', stats.logistic(loc=1.0, scale=2.0)),
        ('Rayleigh(2., 1.)', stats.rayleigh(loc=1.0, scale=2.0)),
    ]
    for text, dist in cases:
        d = distributions.parse(text)
        np.testing.assert_allclose(distributions.inverse_cdf_host(d[0], d[1:], u), dist.ppf(u), rtol=1e-10, atol=1e-12,
                                   err_msg=text)
    # the form every shipped Channel_Flow / Mascaret case uses: mean and standard deviation on [a, b]
    d = distributions.parse('BetaMuSigma(4031, 400, 1000, 6000).getDistribution()')
    assert d[0] == distributions.BETA
    ref = stats.beta(d[1], d[2], loc=d[3], scale=d[4] - d[3])
    assert ref.mean() == pytest.approx(4031.0, rel=1e-12) and ref.std() == pytest.approx(400.0, rel=1e-12)
    d = distributions.parse('LogNormalMuSigma(3., 1.5, 0.5).getDistribution()')
    ref = stats.lognorm(s=d[2], loc=d[3], scale=np.exp(d[1]))
    assert ref.mean() == pytest.approx(3.0, rel=1e-12) and ref.std() == pytest.approx(1.5, rel=1e-12)


def test_minmax_is_sklearn_minmaxscaler():
    from sklearn.preprocessing import MinMaxScaler
    corners = np.array([[15.0, 2500.0, -np.pi], [60.0, 6000.0, np.pi]])
    ref = MinMaxScaler().fit(corners)
    mine = MinMax(corners)
    assert np.array_equal(ref.scale_, mine.scale_) and np.array_equal(ref.min_, mine.min_)
    pts = np.random.RandomState(0).rand(50, 3) * 100
    assert np.array_equal(ref.transform(pts), mine.transform(pts))


def test_no_product_import_of_oracle():
    """The product path must not route through the oracle."""
    pkg = os.path.join(ROOT, 'batman_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract: exactly one JSON line on stdout (everything else on stderr), with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--cpu-sample', '2000'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['value'] > 0 and d['e2e']['h2d_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and 'sobol' in d['cpu_baseline']


def test_lockstep_batcher_batches_all_optimisers_deterministically():
    """fit.LockstepBatcher: R scipy DE runs in threads, ONE objective batch per generation; the result of each run
    equals the same run on its own (the batch composition does not depend on thread timing)."""
    import threading
    from scipy.optimize import differential_evolution
    from batman_b200.fit import LockstepBatcher
    bounds = np.array([[-3.0, 3.0]] * 3)
    centres = {c: np.array([0.5 * c, -0.25 * c, 0.1]) for c in range(5)}

    def f(cid, pts):                                   # pts (S, d) -> (S,)
        return ((pts - centres[cid]) ** 2).sum(axis=1) + 0.1 * np.sin(5 * pts).sum(axis=1)

    sizes = []

    def evaluate(batch):
        sizes.append(sum(len(t) for _, t in batch))
        assert [c for c, _ in batch] == sorted(c for c, _ in batch)
        return [f(cid, t) for cid, t in batch]

    def run(cid, fun):
        return differential_evolution(lambda th: fun(th.T), bounds, tol=1e-3, popsize=8 + cid, vectorized=True,
                                      updating='deferred', polish=False, seed=100 + cid, maxiter=40)

    batcher = LockstepBatcher(evaluate, list(centres))
    results = {}

    def work(cid):
        try:
            results[cid] = run(cid, lambda pts: batcher.submit(cid, pts))
        finally:
            batcher.retire(cid)
    threads = [threading.Thread(target=work, args=(c,)) for c in centres]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert len(results) == 5
    for cid in centres:
        alone = run(cid, lambda pts: f(cid, pts))
        np.testing.assert_array_equal(results[cid].x, alone.x)
        assert results[cid].nfev == alone.nfev
    # the first batches carry all five populations: 3 * (8+9+10+11+12) candidates
    assert sizes[0] == 3 * sum(8 + c for c in centres)
    assert batcher.n_batches == len(sizes) and batcher.n_evals == sum(sizes)
    assert batcher.n_batches <= max(r.nit for r in results.values()) + 2


def test_lockstep_batcher_propagates_failures_without_deadlock():
    import threading
    from batman_b200.fit import LockstepBatcher

    def evaluate(batch):
        raise ValueError("device fault")
    batcher = LockstepBatcher(evaluate, [0, 1])
    errs = []

    def work(cid):
        try:
            batcher.submit(cid, np.zeros((2, 2)))
        except RuntimeError as exc:
            errs.append(exc)
        finally:
            batcher.retire(cid)
    threads = [threading.Thread(target=work, args=(c,)) for c in (0, 1)]
    [t.start() for t in threads]
    [t.join(timeout=20) for t in threads]
    assert len(errs) == 2 and all(isinstance(e.__cause__, ValueError) for e in errs)


def test_sigma_kernel_gang_size_policy():
    """CTAs per tile of test points in the int8 sigma kernel: the K* tiles in flight (SMs / G tiles of 896 n_pad bytes)
    must fit in about half of the 126 MB L2 of a B200 (148 SMs), and G never exceeds the SM count."""
    from batman_b200 import _lib
    lib = _lib.load()
    want = {100: 2, 1024: 2, 2000: 4, 4096: 8, 8192: 16, 16384: 37}
    for n, g in want.items():
        assert lib.bk_sigma_gang_size(n, 148) == g, n
        n_pad = lib.bk_gp_padded_n(n)
        assert (148 // g) * n_pad * 896 <= 72e6
    assert lib.bk_sigma_gang_size(16384, 16) <= 16
    assert lib.bk_sigma_gang_size(0, 148) == 0
