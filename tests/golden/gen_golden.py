"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

Every fixture holds (a) the fitted state scikit-learn produced inside the
reference's own ``batman.surrogate.Kriging`` / ``SurrogateModel`` / ``Pod``
objects and (b) the outputs of the reference's own ``evaluate`` / ``__call__``
/ ``inverse_transform`` on fixed points.  The committed .npz files are what the
CPU and GPU parity tests load; this script is committed so they can be
regenerated.  ``global_optimizer=False`` (L-BFGS-B, seeded) is used because the
default differential-evolution fit is unseeded in the reference
(kriging.py:170-171) and takes minutes.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import functions as ofun  # noqa: E402
from oracle import gp as ogp  # noqa: E402
from oracle import ref_shim  # noqa: E402

warnings.simplefilter("ignore")
ref_shim.import_batman()
from batman.functions import Channel_Flow, G_Function, Ishigami, Michalewicz  # noqa: E402
from batman.pod import Pod  # noqa: E402
from batman.surrogate import Kriging, SurrogateModel  # noqa: E402
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern  # noqa: E402


def tree_to_json(tree):
    if tree[0] in ('sum', 'prod'):
        return [tree[0], tree_to_json(tree[1]), tree_to_json(tree[2])]
    return [tree[0]] + [np.asarray(t).tolist() if isinstance(t, np.ndarray) else t for t in tree[1:]]


def dump_states(gp_models):
    out = {}
    trees = []
    for q, gp in enumerate(gp_models):
        st = ogp.state_from_sklearn(gp)
        trees.append(tree_to_json(st['tree']))
        out['alpha_%d' % q] = st['alpha']
        out['L_%d' % q] = st['L']
        out['ystats_%d' % q] = np.array([st['y_mean'], st['y_std']])
        out['theta_%d' % q] = np.asarray(gp.kernel_.theta)
    out['X_train'] = np.asarray(gp_models[0].X_train_, dtype=np.float64)
    out['trees'] = np.array(json.dumps(trees))
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def case_docstring():
    """kriging.py:14-19 known answer."""
    np.random.seed(123456)
    sample = np.array([[2, 4], [3, 5], [6, 9]])
    data = np.array([[12, 1], [10, 2], [9, 4]])
    k = Kriging(sample, data, global_optimizer=False)
    pts = np.array([[5.0, 8.0], [2.0, 4.0], [4.1, 6.3], [0.0, 0.0], [6.0, 9.0]])
    mean, sigma = k.evaluate(pts)
    save('kriging_docstring', points=pts, ref_mean=mean, ref_sigma=sigma, **dump_states(k.gp_models))


def case_kernels():
    """Kernel variants the reference tests exercise (tests/test_models.py:74-78,
    171-172; test_cases/functionnal_test.py:224-228) plus RBF / Matern 0.5, 2.5."""
    rng = np.random.RandomState(123456)
    variants = {
        'iso_matern_noise': dict(kernel=Matern(), noise=0.8, d=3),
        'iso_matern_noise_true': dict(kernel=Matern(), noise=True, d=3),
        'const_plus_matern05_noise': dict(kernel=ConstantKernel() + Matern(length_scale=0.5, nu=0.5), noise=1, d=2),
        'const_times_rbf_ard': dict(kernel=ConstantKernel() * RBF(length_scale=(1.0,) * 4), noise=False, d=4),
        'const_times_matern25_ard': dict(kernel=ConstantKernel() * Matern(length_scale=(1.0,) * 2, nu=2.5), noise=False, d=2),
        'default_3out': dict(kernel=None, noise=False, d=3, m=3),
    }
    for name, v in variants.items():
        d, m = v['d'], v.get('m', 1)
        np.random.seed(123456)
        X = ofun.halton(40, d)
        Y = np.stack([np.sin(3 * X[:, 0] * (q + 1)) + X[:, -1] ** 2 * (q + 1) + 0.05 * rng.randn(40)
                      for q in range(m)], axis=1)
        k = Kriging(X, Y, kernel=v['kernel'], noise=v['noise'], global_optimizer=False)
        pts = np.vstack([rng.rand(48, d), X[:4], X[:4] + 1e-9])
        mean, sigma = k.evaluate(pts)
        save('kernel_' + name, points=pts, ref_mean=mean, ref_sigma=sigma, **dump_states(k.gp_models))


def case_ishigami():
    """cfg1: Ishigami 3-D, n_train=100, through SurrogateModel.__call__ (scaler included)."""
    np.random.seed(123456)
    corners = [[-np.pi] * 3, [np.pi] * 3]
    X = ofun.halton(100, 3, corners)
    y = np.array(Ishigami()(X))
    assert np.allclose(y, ofun.ishigami(X), rtol=0, atol=1e-13)
    s = SurrogateModel('kriging', corners, ['x1', 'x2', 'x3'], global_optimizer=False)
    s.fit(X, y)
    rng = np.random.RandomState(7)
    pts = -np.pi + 2 * np.pi * rng.rand(256, 3)
    mean, sigma = s(pts)
    save('ishigami_n100', corners=np.array(corners), X_raw=X, y=y, points=pts, ref_mean=mean,
         ref_sigma=sigma, **dump_states(s.predictor.gp_models))


def case_pod():
    """tests/test_pod.py:15-24, 109-114 known answer + SurrogateModel-with-POD path."""
    snapshots = np.array([[37., 40., 41., 49., 42., 46., 45., 48.],
                          [40., 43., 47., 46., 41., 46., 45., 48.],
                          [40., 41., 42., 45., 44., 46., 45., 47.]])
    pod = Pod([[1], [8]], 1, 3)
    pod._fit(snapshots.copy())       # Pod.fit's in-place centring trips pandas-3 CoW
    modes = np.array([[4.602, -0.73, -0.592], [1.575, -3.615, 0.121]])
    save('pod_inverse', mean_snapshot=pod.mean_snapshot, U=pod.U, S=pod.S, V=pod.V, modes=modes,
         ref_inverse=pod.inverse_transform(modes), snapshots=snapshots)

    # Channel_Flow (default dx=8000 -> F=5), 2-D input, POD(0.99, 3) + per-mode Kriging
    np.random.seed(123456)
    corners = [[15.0, 2500.0], [60.0, 6000.0]]
    X = ofun.halton(30, 2, corners)
    fun = Channel_Flow()
    Yf = np.array(fun(X))
    assert np.allclose(Yf, ofun.channel_flow(X), rtol=0, atol=1e-11)
    pod = Pod(corners, 0.99, 3)
    pod._fit(Yf.T.copy())
    modes_train = pod.VS                              # what the surrogate is fitted on (driver.py)
    s = SurrogateModel('kriging', corners, ['Ks', 'Q'], global_optimizer=False)
    s.fit(X, modes_train, pod=pod)
    rng = np.random.RandomState(11)
    pts = np.array(corners[0]) + (np.array(corners[1]) - np.array(corners[0])) * rng.rand(64, 2)
    mean, sigma = s(pts)
    save('channel_flow_pod', corners=np.array(corners), X_raw=X, Yf=Yf, mean_snapshot=pod.mean_snapshot,
         U=pod.U, S=pod.S, modes_train=modes_train, points=pts, ref_mean=mean, ref_sigma=sigma,
         **dump_states(s.predictor.gp_models))


def case_functions():
    rng = np.random.RandomState(3)
    x3 = -np.pi + 2 * np.pi * rng.rand(16, 3)
    x8 = rng.rand(16, 8)
    x10 = np.pi * rng.rand(16, 10)
    x2 = np.array([15.0, 2500.0]) + np.array([45.0, 3500.0]) * rng.rand(4, 2)
    save('functions', x3=x3, ishigami=np.array(Ishigami()(x3)), x8=x8, g_function=np.array(G_Function(d=8)(x8)),
         x10=x10, michalewicz=np.array(Michalewicz(d=10)(x10)), x2=x2,
         channel_flow=np.array(Channel_Flow(dx=100., length=40000.)(x2)),
         g_s_first=G_Function(d=8).s_first, g_s_total=G_Function(d=8).s_total,
         ishigami_s_first=Ishigami().s_first, ishigami_s_total2=Ishigami().s_total2,
         ishigami_s_second=Ishigami().s_second)


if __name__ == '__main__':
    case_docstring()
    case_kernels()
    case_ishigami()
    case_pod()
    case_functions()
