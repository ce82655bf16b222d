"""``np.longdouble`` (x87 80-bit, 64-bit mantissa) restatement of oracle/gp.py.

TEST INFRASTRUCTURE ONLY.  Arbiter when scikit-learn and the CUDA path
disagree near the 1e-10 gate: the formulas of sklearn:_gpr.py:444-500 evaluated
with ~11 more mantissa bits on the *same float64 inputs* (X, theta, alpha, L).
Slow (python loop over the n rows of L): keep n <= ~2048 and a few hundred
test points.
"""
import numpy as np

from . import gp

LD = np.longdouble


def predict_hp(state, Xs, return_std=True):
    Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
    K_trans = gp.kernel_cross(state['tree'], Xs, state['X_train'], dtype=LD)   # (k, n) longdouble
    alpha = state['alpha'].astype(LD)
    y_mean = (K_trans * alpha[None, :]).sum(axis=1)
    y_mean = LD(state['y_std']) * y_mean + LD(state['y_mean'])
    if not return_std:
        return y_mean
    L = state['L'].astype(LD)
    n = L.shape[0]
    rhs = K_trans.T.copy()
    V = np.zeros_like(rhs)
    for i in range(n):  # forward substitution, vectorised over test points
        V[i] = (rhs[i] - L[i, :i] @ V[:i]) / L[i, i]
    y_var = LD(gp.kernel_diag(state['tree'])) - (V * V).sum(axis=0)
    y_var = np.where(y_var < 0, LD(0), y_var)
    y_var = y_var * LD(state['y_std']) ** 2
    return y_mean, np.sqrt(y_var)


def amplification(state, Xs):
    """sum|k_j alpha_j| / |sum k_j alpha_j|: how many digits the mean sum eats."""
    K_trans = gp.kernel_cross(state['tree'], np.atleast_2d(Xs), state['X_train'])
    terms = K_trans * state['alpha'][None, :]
    return np.abs(terms).sum(axis=1) / np.maximum(np.abs(terms.sum(axis=1)), 1e-300)
