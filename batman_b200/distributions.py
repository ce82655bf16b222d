"""Marginal distributions of the UQ inputs.

The reference builds ``ot.ComposedDistribution([ot.<string>, ...])`` by
``eval`` of strings such as ``"Uniform(15., 60.)"``, ``"Normal(4035., 400.)"``
or ``"BetaMuSigma(4031, 400, 1000, 6000).getDistribution()"`` (uq/uq.py:151-158;
the three forms the shipped test cases use).  OpenTURNS is not available, so
the strings are parsed here -- no ``eval`` -- with the OpenTURNS 1.15
parametrisations and defaults, and mapped to the device generator: the inverse
CDF of a Philox uniform (csrc/sobol_common.cuh ``marginal``).  An unknown name
raises ``SystemError`` as the reference does (uq.py:156-158); an OpenTURNS
family without a device inverse CDF raises ``NotImplementedError`` -- pass
explicit base samples then.

Every marginal is ``(kind, par0..par3)`` with ``NPAR = 4`` parameters (unused 0).
"""
import math
import re

(UNIFORM, NORMAL, LOGNORMAL, TRIANGULAR, TRUNCNORMAL, BETA, GUMBEL, EXPONENTIAL, LOGUNIFORM, WEIBULLMIN, LOGISTIC,
 RAYLEIGH) = range(12)
NPAR = 4
NAMES = ['uniform', 'normal', 'lognormal', 'triangular', 'truncatednormal', 'beta', 'gumbel', 'exponential',
         'loguniform', 'weibullmin', 'logistic', 'rayleigh']

_OT_FAMILIES = {'Arcsine', 'Beta', 'BetaMuSigma', 'Burr', 'Chi', 'ChiSquare', 'Dirichlet', 'Epanechnikov',
                'Exponential', 'FisherSnedecor', 'Frechet', 'Gamma', 'GammaMuSigma', 'GeneralizedPareto',
                'Gumbel', 'GumbelMuSigma', 'Histogram', 'InverseNormal', 'Laplace', 'Logistic', 'LogNormal',
                'LogNormalMuSigma', 'LogUniform', 'Rayleigh', 'Rice', 'Student', 'Trapezoidal', 'Triangular',
                'TruncatedNormal', 'Weibull', 'WeibullMin', 'WeibullMax', 'Uniform', 'Normal'}

_CALL = re.compile(r'^\s*(?:ot\.)?([A-Za-z_]\w*)\s*\(([^()]*)\)\s*(\.getDistribution\(\s*\))?\s*$')


def _phi(x):
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def _fill(vals, defaults):
    """OpenTURNS constructors take their parameters positionally; missing ones take the defaults."""
    if len(vals) > len(defaults):
        raise SystemError('OpenTURNS distribution unknown.')        # the eval of uq.py:152 raises TypeError -> SystemError
    return list(vals) + list(defaults[len(vals):])


def _bad():
    raise SystemError('OpenTURNS distribution unknown.')


def _build(name, vals, parametrised):
    """-> (kind, p0, p1, p2, p3) or None if the family has no device inverse CDF."""
    if name == 'Uniform':
        a, b = _fill(vals, [-1.0, 1.0])
        if not a < b:
            _bad()
        return (UNIFORM, a, b, 0.0, 0.0)
    if name == 'Normal':
        mu, sigma = _fill(vals, [0.0, 1.0])
        if not sigma > 0:
            _bad()
        return (NORMAL, mu, sigma, 0.0, 0.0)
    if name == 'LogNormal':
        mu_log, sigma_log, gamma = _fill(vals, [0.0, 1.0, 0.0])
        if not sigma_log > 0:
            _bad()
        return (LOGNORMAL, mu_log, sigma_log, gamma, 0.0)
    if name == 'LogNormalMuSigma':                                   # (mu, sigma, gamma) of the variable itself
        mu, sigma, gamma = _fill(vals, [math.exp(0.5), math.sqrt(math.exp(2.0) - math.exp(1.0)), 0.0])
        if not (sigma > 0 and mu > gamma):
            _bad()
        s2 = math.log1p((sigma / (mu - gamma)) ** 2)
        return (LOGNORMAL, math.log(mu - gamma) - 0.5 * s2, math.sqrt(s2), gamma, 0.0)
    if name == 'Triangular':
        a, m, b = _fill(vals, [-1.0, 0.0, 1.0])
        if not (a <= m <= b and a < b):
            _bad()
        return (TRIANGULAR, a, m, b, 0.0)
    if name == 'TruncatedNormal':
        mu, sigma, a, b = _fill(vals, [0.0, 1.0, -1.0, 1.0])
        if not (sigma > 0 and a < b):
            _bad()
        return (TRUNCNORMAL, mu, sigma, _phi((a - mu) / sigma), _phi((b - mu) / sigma))
    if name == 'Beta':                                               # OpenTURNS >= 1.14: (alpha, beta, a, b)
        alpha, beta, a, b = _fill(vals, [2.0, 2.0, -1.0, 1.0])
        if not (alpha > 0 and beta > 0 and a < b):
            _bad()
        return (BETA, alpha, beta, a, b)
    if name == 'BetaMuSigma':                                        # mean / standard deviation on [a, b]
        mu, sigma, a, b = _fill(vals, [0.5, 1.0 / math.sqrt(12.0) * 0.5, 0.0, 1.0])
        if not (sigma > 0 and a < mu < b):
            _bad()
        t = (b - mu) * (mu - a) / sigma ** 2 - 1.0
        if not t > 0:
            _bad()
        alpha = t * (mu - a) / (b - a)
        return (BETA, alpha, t - alpha, a, b)
    if name == 'Gumbel':                                             # OpenTURNS >= 1.14: (beta, gamma)
        beta, gamma = _fill(vals, [1.0, 0.0])
        if not beta > 0:
            _bad()
        return (GUMBEL, beta, gamma, 0.0, 0.0)
    if name == 'Exponential':
        lam, gamma = _fill(vals, [1.0, 0.0])
        if not lam > 0:
            _bad()
        return (EXPONENTIAL, lam, gamma, 0.0, 0.0)
    if name == 'LogUniform':
        a_log, b_log = _fill(vals, [-1.0, 1.0])
        if not a_log < b_log:
            _bad()
        return (LOGUNIFORM, a_log, b_log, 0.0, 0.0)
    if name in ('WeibullMin', 'Weibull'):                            # (beta, alpha, gamma); Weibull = deprecated alias
        beta, alpha, gamma = _fill(vals, [1.0, 1.0, 0.0])
        if not (beta > 0 and alpha > 0):
            _bad()
        return (WEIBULLMIN, beta, alpha, gamma, 0.0)
    if name == 'Logistic':
        mu, beta = _fill(vals, [0.0, 1.0])
        if not beta > 0:
            _bad()
        return (LOGISTIC, mu, beta, 0.0, 0.0)
    if name == 'Rayleigh':
        beta, gamma = _fill(vals, [1.0, 0.0])
        if not beta > 0:
            _bad()
        return (RAYLEIGH, beta, gamma, 0.0, 0.0)
    return None


_PARAMETRISED = {'BetaMuSigma', 'LogNormalMuSigma', 'GumbelMuSigma', 'GammaMuSigma'}


def parse(dist):
    """'Uniform(a, b)' | 'BetaMuSigma(mu, sigma, a, b).getDistribution()' | ('uniform', a, b) | ... ->
    (kind, par0, par1, par2, par3)."""
    if isinstance(dist, (tuple, list)) and len(dist) >= 1 and str(dist[0]).lower() in NAMES:
        kind = NAMES.index(str(dist[0]).lower())
        vals = [float(v) for v in dist[1:]]
        if len(vals) > NPAR:
            _bad()
        return tuple([kind] + vals + [0.0] * (NPAR - len(vals)))
    if isinstance(dist, str):
        m = _CALL.match(dist)
        if m:
            name, args, getter = m.group(1), m.group(2), m.group(3)
            try:
                vals = [float(a) for a in args.split(',')] if args.strip() else []
            except ValueError:
                vals = None
            if name in _OT_FAMILIES and vals is not None:
                # a *MuSigma object is a parameter set: the reference's settings spell out .getDistribution()
                if (name in _PARAMETRISED) != bool(getter):
                    _bad()
                built = _build(name, vals, bool(getter))
                if built is not None:
                    return built
                raise NotImplementedError(
                    "distribution %r has no device inverse-CDF; pass base_samples=(A, B) drawn on the host" % dist)
    raise SystemError('OpenTURNS distribution unknown.')


def parse_all(dists):
    """-> (kinds (p,), params (NPAR * p,): par_k of marginal c at NPAR * c + k)."""
    parsed = [parse(d) for d in dists]
    kinds = [d[0] for d in parsed]
    params = [v for d in parsed for v in d[1:]]
    return kinds, params


def inverse_cdf_host(kind, par, u):
    """numpy/scipy statement of the device inverse CDFs (the LHS propagation sample of UQ uses it only when no
    surrogate -- hence no device -- is involved)."""
    import numpy as np
    from scipy import special
    a, b, c, d = par
    u = np.asarray(u, dtype=np.float64)
    if kind == UNIFORM:
        return a + (b - a) * u
    if kind == NORMAL:
        return a + b * special.ndtri(u)
    if kind == LOGNORMAL:
        return c + np.exp(a + b * special.ndtri(u))
    if kind == TRIANGULAR:
        fm = (b - a) / (c - a)
        return np.where(u < fm, a + np.sqrt(u * (c - a) * (b - a)), c - np.sqrt((1.0 - u) * (c - a) * (c - b)))
    if kind == TRUNCNORMAL:
        return a + b * special.ndtri(c + u * (d - c))
    if kind == BETA:
        return c + (d - c) * special.betaincinv(a, b, u)
    if kind == GUMBEL:
        return b - a * np.log(-np.log(u))
    if kind == EXPONENTIAL:
        return b - np.log1p(-u) / a
    if kind == LOGUNIFORM:
        return np.exp(a + (b - a) * u)
    if kind == WEIBULLMIN:
        return c + a * np.power(-np.log1p(-u), 1.0 / b)
    if kind == LOGISTIC:
        return a + b * (np.log(u) - np.log1p(-u))
    if kind == RAYLEIGH:
        return b + a * np.sqrt(-2.0 * np.log1p(-u))
    raise ValueError("unknown marginal kind %r" % (kind,))
