"""Marginal distributions of the UQ inputs.

The reference builds ``ot.ComposedDistribution([ot.<string>, ...])`` by
``eval`` of strings such as ``"Uniform(15., 60.)"`` or ``"Normal(4035., 400.)"``
(uq/uq.py:151-158).  OpenTURNS is not available; the two families every shipped
test case of this path uses are parsed here and mapped to the device generator
(inverse-CDF of a Philox uniform).  An unknown name raises ``SystemError`` as
the reference does (uq.py:156-158); a known OpenTURNS family without a device
inverse-CDF raises ``NotImplementedError`` -- pass explicit base samples then.
"""
import re

UNIFORM, NORMAL = 0, 1

_OT_FAMILIES = {'Arcsine', 'Beta', 'BetaMuSigma', 'Burr', 'Chi', 'ChiSquare', 'Dirichlet', 'Epanechnikov',
                'Exponential', 'FisherSnedecor', 'Frechet', 'Gamma', 'GammaMuSigma', 'GeneralizedPareto',
                'Gumbel', 'GumbelMuSigma', 'Histogram', 'InverseNormal', 'Laplace', 'Logistic', 'LogNormal',
                'LogNormalMuSigma', 'LogUniform', 'Rayleigh', 'Rice', 'Student', 'Trapezoidal', 'Triangular',
                'TruncatedNormal', 'Weibull', 'WeibullMin', 'WeibullMax'}

_CALL = re.compile(r'^\s*(?:ot\.)?([A-Za-z_]\w*)\s*\(([^()]*)\)\s*(.*)$')


def parse(dist):
    """'Uniform(a, b)' | 'Normal(mu, sigma)' | ('uniform', a, b) -> (kind, par0, par1)."""
    if isinstance(dist, (tuple, list)) and len(dist) == 3:
        name = str(dist[0]).lower()
        if name in ('uniform', 'normal'):
            return (UNIFORM if name == 'uniform' else NORMAL, float(dist[1]), float(dist[2]))
    if isinstance(dist, str):
        m = _CALL.match(dist)
        if m:
            name, args, tail = m.group(1), m.group(2), m.group(3)
            try:
                vals = [float(a) for a in args.split(',')] if args.strip() else []
            except ValueError:
                vals = None
            if name == 'Uniform' and vals is not None and not tail:
                a, b = (vals + [-1.0, 1.0][len(vals):])[:2] if len(vals) < 2 else vals[:2]
                if not a < b:
                    raise SystemError('OpenTURNS distribution unknown.')
                return (UNIFORM, a, b)
            if name == 'Normal' and vals is not None and not tail:
                mu, sigma = (vals + [0.0, 1.0][len(vals):])[:2] if len(vals) < 2 else vals[:2]
                if not sigma > 0:
                    raise SystemError('OpenTURNS distribution unknown.')
                return (NORMAL, mu, sigma)
            if name in _OT_FAMILIES or name in ('Uniform', 'Normal'):
                raise NotImplementedError(
                    "distribution %r has no device inverse-CDF; pass base_samples=(A, B) drawn on the host" % dist)
    raise SystemError('OpenTURNS distribution unknown.')


def parse_all(dists):
    parsed = [parse(d) for d in dists]
    kinds = [d[0] for d in parsed]
    params = [v for d in parsed for v in d[1:]]
    return kinds, params
