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
