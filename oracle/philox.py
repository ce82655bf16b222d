"""Counter-based Saltelli base-sample generator, numpy side.

TEST INFRASTRUCTURE ONLY.  The reference draws A and B with OpenTURNS'
``SobolIndicesExperiment`` (uq/uq.py:331-332; dSFMT stream, not reproducible
without OpenTURNS).  For N*(2p+2) up to 2.2e9 rows the design cannot be stored,
so the CUDA path generates A_k / B_k in-kernel from a counter-based generator.
This file is the bit-exact CPU statement of that generator (Philox4x32-10,
Salmon et al. 2011, public algorithm), so "same inputs and seeds" is defined:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (row & 0xffffffff, row >> 32, col // 2, stream)   stream 0 = A, 1 = B
    words   = philox4x32_10(counter, key)           -> w0..w3
    column 2*(col//2)   uses (w0, w1); column 2*(col//2)+1 uses (w2, w3)
    u       = (((hi >> 6) << 26 | (lo >> 6)) + 0.5) * 2**-52      in (0, 1), exact
    Uniform(a, b): x = a + (b - a) * u        (two roundings, no FMA)
    Normal(mu, s): x = mu + s * ndtri(u)      (ndtri differs by ulps between
                                               libraries: parity tests that need
                                               identical inputs dump the device
                                               design instead)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint64(0x9E3779B9)
W1 = np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> S32, p0 & MASK
        hi1, lo1 = p1 >> S32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def _to_unit(hi, lo):
    k = ((hi >> np.uint64(6)) << np.uint64(26)) | (lo >> np.uint64(6))
    return (k.astype(np.float64) + 0.5) * 2.0 ** -52


def uniforms(seed, rows, p, stream):
    """u[row, col] for the given absolute row indices: shape (len(rows), p)."""
    rows = np.asarray(rows, dtype=np.uint64)
    out = np.empty((len(rows), p), dtype=np.float64)
    k0 = np.uint64(seed) & MASK
    k1 = (np.uint64(seed) >> S32) & MASK
    for cp in range((p + 1) // 2):
        w0, w1, w2, w3 = philox4x32_10(rows & MASK, rows >> S32,
                                       np.full_like(rows, cp), np.full_like(rows, stream), k0, k1)
        out[:, 2 * cp] = _to_unit(w0, w1)
        if 2 * cp + 1 < p:
            out[:, 2 * cp + 1] = _to_unit(w2, w3)
    return out


def marginal_transform(u, dists):
    """dists: one tuple per column, ('uniform', a, b) | ('normal', mu, sigma) | ('lognormal', muLog, sigmaLog, gamma) |
    ('triangular', a, m, b) | ('truncatednormal', mu, sigma, Phi_a, Phi_b) | ('beta', alpha, beta, a, b) |
    ('gumbel', beta, gamma) | ('exponential', lambda, gamma) | ('loguniform', aLog, bLog) |
    ('weibullmin', beta, alpha, gamma) | ('logistic', mu, beta) | ('rayleigh', beta, gamma): the inverse CDFs of
    OpenTURNS 1.15's parametrisations on scipy.special (the device twins: csrc/sobol_common.cuh ``marginal``)."""
    from scipy import special
    x = np.empty_like(u)
    for j, d in enumerate(dists):
        name = d[0]
        par = [float(v) for v in d[1:]] + [0.0] * (5 - len(d))
        a, b, c, e = par
        uj = u[:, j]
        if name == 'uniform':
            x[:, j] = a + (b - a) * uj
        elif name == 'normal':
            x[:, j] = a + b * special.ndtri(uj)
        elif name == 'lognormal':
            x[:, j] = c + np.exp(a + b * special.ndtri(uj))
        elif name == 'triangular':
            fm = (b - a) / (c - a)
            x[:, j] = np.where(uj < fm, a + np.sqrt(uj * (c - a) * (b - a)), c - np.sqrt((1.0 - uj) * (c - a) * (c - b)))
        elif name == 'truncatednormal':
            x[:, j] = a + b * special.ndtri(c + uj * (e - c))
        elif name == 'beta':
            x[:, j] = c + (e - c) * special.betaincinv(a, b, uj)
        elif name == 'gumbel':
            x[:, j] = b - a * np.log(-np.log(uj))
        elif name == 'exponential':
            x[:, j] = b - np.log1p(-uj) / a
        elif name == 'loguniform':
            x[:, j] = np.exp(a + (b - a) * uj)
        elif name == 'weibullmin':
            x[:, j] = c + a * np.power(-np.log1p(-uj), 1.0 / b)
        elif name == 'logistic':
            x[:, j] = a + b * (np.log(uj) - np.log1p(-uj))
        elif name == 'rayleigh':
            x[:, j] = b + a * np.sqrt(-2.0 * np.log1p(-uj))
        else:
            raise NotImplementedError(name)
    return x


def base_samples(seed, row_lo, row_hi, dists):
    """A[row_lo:row_hi], B[row_lo:row_hi] of the Saltelli experiment."""
    rows = np.arange(row_lo, row_hi, dtype=np.uint64)
    p = len(dists)
    A = marginal_transform(uniforms(seed, rows, p, 0), dists)
    B = marginal_transform(uniforms(seed, rows, p, 1), dists)
    return A, B


def lhs_sample(seed, perm, dists):
    """Randomised LHS (ot.LHSExperiment(dist, N, True, True), uq/uq.py:159-160) as the device draws it
    (csrc/uq_extra.cu k_lhs_rows): cell perm[i, c] of row i in dimension c, in-cell shift = Philox stream 2,
    u = (cell + shift) / N, then the marginal's inverse CDF.  perm: (N, p) integer array, every column a permutation."""
    perm = np.asarray(perm)
    n, p = perm.shape
    shift = uniforms(seed, np.arange(n), p, 2)
    return marginal_transform((perm.astype(np.float64) + shift) * (1.0 / n), dists)
