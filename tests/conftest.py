import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Safety net for a checkout without the built library (it is git-ignored): compile it once with nvcc.  The
    # product itself never does this -- batman_b200._lib.load() raises ImportError when the .so is missing.
    lib = os.path.join(ROOT, 'batman_b200', 'libbatman_b200.so')
    if not os.path.exists(lib):
        sys.stderr.write("tests: %s is missing, building it with nvcc (about two minutes)...\n" % lib)
        from batman_b200.csrc import build as _build
        _build.build(force=True)


def _tree_from_json(node):
    if node[0] in ('sum', 'prod'):
        return (node[0], _tree_from_json(node[1]), _tree_from_json(node[2]))
    if node[0] == 'matern':
        return ('matern', float(node[1]), np.atleast_1d(np.asarray(node[2], dtype=np.float64)))
    if node[0] == 'rbf':
        return ('rbf', np.atleast_1d(np.asarray(node[1], dtype=np.float64)))
    return (node[0], float(node[1]))


def load_golden(name):
    """-> (npz dict, [oracle state per output])."""
    z = dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))
    states = []
    if 'trees' in z:
        trees = json.loads(str(z['trees']))
        for q, t in enumerate(trees):
            states.append(dict(tree=_tree_from_json(t), X_train=z['X_train'], alpha=z['alpha_%d' % q],
                               L=z['L_%d' % q], y_mean=float(z['ystats_%d' % q][0]),
                               y_std=float(z['ystats_%d' % q][1])))
    return z, states


GP_CASES = ['kriging_docstring', 'kernel_iso_matern_noise', 'kernel_iso_matern_noise_true',
            'kernel_const_plus_matern05_noise', 'kernel_const_times_rbf_ard',
            'kernel_const_times_matern25_ard', 'kernel_default_3out']


@pytest.fixture(scope='session')
def golden():
    return load_golden


def tree_to_sklearn(tree):
    """oracle kernel tree -> sklearn kernel object (what the reference hands to Kriging)."""
    from sklearn.gaussian_process import kernels as sk
    tag = tree[0]
    if tag == 'const':
        return sk.ConstantKernel(tree[1])
    if tag == 'white':
        return sk.WhiteKernel(tree[1])
    if tag == 'matern':
        ls = tree[2]
        return sk.Matern(length_scale=float(ls[0]) if len(ls) == 1 else ls, nu=tree[1])
    if tag == 'rbf':
        ls = tree[1]
        return sk.RBF(length_scale=float(ls[0]) if len(ls) == 1 else ls)
    if tag == 'sum':
        return sk.Sum(tree_to_sklearn(tree[1]), tree_to_sklearn(tree[2]))
    return sk.Product(tree_to_sklearn(tree[1]), tree_to_sklearn(tree[2]))


class StateGP:
    """Duck of a fitted GaussianProcessRegressor built from an oracle state."""

    def __init__(self, st):
        self.kernel_ = tree_to_sklearn(st['tree'])
        self.X_train_ = st['X_train']
        self.alpha_ = st['alpha']
        self.L_ = st['L']
        self._y_train_mean = st['y_mean']
        self._y_train_std = st['y_std']


def kriging_from_states(states):
    from batman_b200 import Kriging
    return Kriging.from_state([StateGP(s) for s in states])


def synthetic_state(n, p, kind='matern32', ls=None, seed=0, func=None, c=1.5):
    """Deterministic fixed-theta model (no optimiser): Halton design, G-function data."""
    from oracle import functions as ofun
    from oracle import gp as ogp
    X = ofun.halton(n, p)
    y = (func or ofun.g_function)(X)[:, 0]
    ls = np.full(p, 0.7) if ls is None else np.asarray(ls, dtype=np.float64)
    stat = {'matern32': ('matern', 1.5, ls), 'matern52': ('matern', 2.5, ls), 'matern12': ('matern', 0.5, ls),
            'rbf': ('rbf', ls)}[kind]
    tree = ('prod', ('const', c), stat)
    return ogp.fit_state(tree, X, y)
