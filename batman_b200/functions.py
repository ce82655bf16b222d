"""The analytical functions ``UQ(test=...)`` names (uq/uq.py:220-284, functions/analytical.py).

``UQ.error_model`` compares the surrogate with a closed-form function: Q2 and
MSE on the propagation sample, L2 distance of the Sobol' indices to the
closed-form ones.  Only the two functions whose indices the reference's own
UQ tests use are restated here (``Ishigami`` analytical.py:260-311,
``G_Function`` :314-361), whole-array instead of one point per call; any object
with ``__call__``, ``s_first``, ``s_second`` and ``s_total`` can be passed to
``UQ(test=...)`` instead of a name (e.g. a ``batman.functions`` instance).
"""
import numpy as np


class Ishigami:
    r"""F = sin(x_1) + a sin(x_2)^2 + b x_3^4 sin(x_1),  x in [-pi, pi]^3."""

    def __init__(self, a=7., b=0.1):
        self.d_in, self.d_out = 3, 1
        self.a, self.b = a, b
        var = 0.5 + a ** 2 / 8 + b * np.pi ** 4 / 5 + b ** 2 * np.pi ** 8 / 18
        v1 = 0.5 + b * np.pi ** 4 / 5 + b ** 2 * np.pi ** 8 / 50
        v2 = a ** 2 / 8
        v13 = b ** 2 * np.pi ** 8 * 8 / 225
        self.s_first = np.array([v1 / var, v2 / var, 0.])
        self.s_second = np.array([[0., 0., v13 / var], [0., 0., 0.], [v13 / var, 0., 0.]])
        self.s_total2 = self.s_first + self.s_second.sum(axis=1)
        self.s_total = np.array([0.558, 0.442, 0.244])              # the rounded values of analytical.py:297

    def __call__(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        return (np.sin(x[:, 0]) + self.a * np.sin(x[:, 1]) ** 2 + self.b * x[:, 2] ** 4 * np.sin(x[:, 0]))[:, None]


class G_Function:
    r"""F = prod_i (|4 x_i - 2| + a_i) / (1 + a_i),  x in [0, 1]^d."""

    def __init__(self, d=4, a=None):
        self.d_in, self.d_out = d, 1
        self.a = np.arange(1, d + 1) if a is None else np.array(a)
        vi = 1. / (3 * (1 + self.a) ** 2)
        v = -1 + np.prod(1 + vi)
        self.s_first = vi / v
        self.s_second = np.zeros((d, d))
        self.s_total = vi * np.prod(1 + vi) / v                     # as analytical.py:345

    def __call__(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        f = np.ones(len(x))
        for i in range(self.d_in):
            f = f * ((np.abs(4. * x[:, i] - 2) + self.a[i]) / (1. + self.a[i]))
        return f[:, None]


def resolve(function):
    """Name (uq.py:245 ``func_ref.__dict__[function]()``) or ready-made object -> function object."""
    if not isinstance(function, str):
        for attr in ('s_first', 's_second', 's_total'):
            if not hasattr(function, attr):
                raise TypeError("UQ(test=...): the test function object needs %r" % attr)
        return function
    table = {'Ishigami': Ishigami, 'G_Function': G_Function}
    if function in table:
        return table[function]()
    try:                                    # a process that has the reference installed can use all of them
        from batman.functions import analytical
        return analytical.__dict__[function]()
    except (ImportError, KeyError):
        raise NotImplementedError(
            "UQ(test=%r): only %s are built in; pass a function object with s_first, s_second, s_total "
            "(e.g. a batman.functions instance)" % (function, sorted(table)))
