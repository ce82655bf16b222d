"""Vectorised analytical test functions used as training-data generators.

TEST INFRASTRUCTURE ONLY.  Formulas from the docstrings of
functions/analytical.py (Ishigami :262, G_Function :317, Michalewicz :143,
Channel_Flow :411-414); the reference evaluates them one point at a time
through ``multi_eval`` -- these are whole-array numpy versions, pinned against
the reference in tests/golden (functions.npz).
"""
import numpy as np


def ishigami(X, a=7.0, b=0.1):
    X = np.atleast_2d(X)
    return (np.sin(X[:, 0]) + a * np.sin(X[:, 1]) ** 2 + b * X[:, 2] ** 4 * np.sin(X[:, 0]))[:, None]


def ishigami_indices(a=7.0, b=0.1):
    """Closed forms, functions/analytical.py:283-296 (s_total2, not the rounded s_total)."""
    pi = np.pi
    var = 0.5 + a ** 2 / 8 + b * pi ** 4 / 5 + b ** 2 * pi ** 8 / 18
    v1 = 0.5 + b * pi ** 4 / 5 + b ** 2 * pi ** 8 / 50
    v2 = a ** 2 / 8
    v13 = b ** 2 * pi ** 8 * 8 / 225
    s1 = np.array([v1, v2, 0.0]) / var
    s2 = np.zeros((3, 3))
    s2[0, 2] = s2[2, 0] = v13 / var
    return s1, s1 + s2.sum(axis=1), s2


def g_function(X, a=None):
    X = np.atleast_2d(X)
    d = X.shape[1]
    a = np.arange(1, d + 1, dtype=np.float64) if a is None else np.asarray(a, dtype=np.float64)
    f = np.ones(len(X))
    for i in range(d):  # same left-to-right product order as analytical.py:357-359
        f = f * ((np.abs(4.0 * X[:, i] - 2) + a[i]) / (1.0 + a[i]))
    return f[:, None]


def g_function_indices(d, a=None):
    """functions/analytical.py:341-345."""
    a = np.arange(1, d + 1, dtype=np.float64) if a is None else np.asarray(a, dtype=np.float64)
    vi = 1.0 / (3 * (1 + a) ** 2)
    v = -1 + np.prod(1 + vi)
    return vi / v, vi * np.prod(1 + vi) / v


def g_function_total_exact(d, a=None):
    """Exact total indices V_i * prod_{j != i}(1 + V_j) / V.  The reference's
    ``s_total`` (analytical.py:345) multiplies over ALL j, i.e. is too large by
    (1 + V_i); the Monte-Carlo anchor uses this exact form instead."""
    a = np.arange(1, d + 1, dtype=np.float64) if a is None else np.asarray(a, dtype=np.float64)
    vi = 1.0 / (3 * (1 + a) ** 2)
    v = -1 + np.prod(1 + vi)
    return vi * np.prod(1 + vi) / (1 + vi) / v


def michalewicz(X, m=10):
    X = np.atleast_2d(X)
    f = np.zeros(len(X))
    for i in range(X.shape[1]):
        f = f + np.sin(X[:, i]) * np.sin((i + 1) * X[:, i] ** 2 / np.pi) ** (2 * m)
    return (-f)[:, None]


def channel_flow(X, dx=8000.0, length=40000.0, width=500.0, slope=5e-4, hinit=10.0):
    """Backward Euler sweep of the backwater curve, vectorised over the sample."""
    X = np.atleast_2d(X)
    ks, q = X[:, 0], X[:, 1]
    g = 9.8
    xs = np.arange(dx, length + 1, dx)
    dl = int(length // dx)
    zref = -xs * slope
    hc = np.power(q ** 2 / (g * width ** 2), 1.0 / 3.0)
    hn = np.power(q ** 2 / (slope * width ** 2 * ks ** 2), 3.0 / 10.0)
    h = hinit * np.ones((len(X), dl))
    for i in range(2, dl + 1):
        prev = h[:, dl - i + 1]
        h[:, dl - i] = prev - dx * slope * ((1 - np.power(prev / hn, -10.0 / 3.0))
                                            / (1 - np.power(prev / hc, -3.0)))
    return zref[None, :] + h


def halton(n, p, corners=None):
    """Unscrambled Halton points (the reference's default DoE is OT's Halton,
    settings "method": "halton"; OT is absent so scipy's is used and recorded)."""
    from scipy.stats import qmc
    pts = qmc.Halton(p, scramble=False).random(n + 1)[1:]
    if corners is not None:
        lo, hi = np.asarray(corners[0], float), np.asarray(corners[1], float)
        pts = lo + (hi - lo) * pts
    return pts
