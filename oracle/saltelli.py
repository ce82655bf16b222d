"""numpy restatement of the OpenTURNS Saltelli design + estimator used by ``UQ.sobol``.

TEST INFRASTRUCTURE ONLY.

**PARITY UNPINNED.**  The reference calls ``ot.SobolIndicesExperiment`` and
``ot.SaltelliSensitivityAlgorithm`` (uq/uq.py:331-332, 342-349, 367-375,
423-426) from ``openturns==1.15`` (requirements.txt:29), a third-party C++
library that is neither vendored under /root/reference nor installable in this
image (no network).  What follows restates the published algorithm
(OpenTURNS 1.15 ``SobolIndicesExperiment::generate``,
``SobolIndicesAlgorithmImplementation`` and
``SaltelliSensitivityAlgorithm::computeIndices``; Saltelli 2002,
"Making best use of model evaluations to compute sensitivity indices").  The
conventions frozen here are:

* design blocks  ``[A; B; E^1..E^p; C^1..C^p]`` with ``E^i`` = A with column i
  from B, ``C^i`` = B with column i from A; the C blocks exist only for second
  order and p > 2 (for p == 2, C^1 == E^2 and C^2 == E^1);
* centring       every output column is centred by its mean over ALL R rows
  before any product is formed;
* ``Var``        = var(yA, ddof=1)  (centring-invariant);
* ``f0^2``       = sum(yA*yB) / N;   ``muA`` = mean(yA) of the centred block;
* first order    ``V_i  = sum(yB*yE^i)/(N-1) - f0^2``;
* total order    ``VT_i = Var - sum(yA*yE^i)/(N-1) + muA^2``;
* second order   ``S_ij = (sum(yE^i*yC^j)/(N-1) - f0^2)/Var - S_i - S_j``;
* aggregated     ``sum_q V_i^q / sum_q Var^q`` (likewise total).

They are anchored only on the analytic indices of Ishigami / G-function
(functions/analytical.py:283-296, 341-345) within Monte-Carlo error, never on
OpenTURNS output.  batman's own python glue (uq/uq.py:400-436, 467-471) IS
restated from source and is pinned by construction.
"""
import numpy as np

LD = np.longdouble


def n_blocks(p, second_order=True):
    """Rows per base point: 2p+2, or p+2 for first/total only or p == 2 (uq.py:89)."""
    return 2 * p + 2 if (second_order and p > 2) else p + 2


def design_from_base(A, B, second_order=True):
    """SobolIndicesExperiment.generate() layout, block-major (R, p)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    p = A.shape[1]
    blocks = [A, B]
    for i in range(p):
        E = A.copy()
        E[:, i] = B[:, i]
        blocks.append(E)
    if second_order and p > 2:
        for i in range(p):
            C = B.copy()
            C[:, i] = A[:, i]
            blocks.append(C)
    return np.concatenate(blocks, axis=0)


def _dot(a, b):
    return (a.astype(LD) * b.astype(LD)).sum(axis=0)


def saltelli_indices(Y, N, p, second_order=True):
    """SaltelliSensitivityAlgorithm(X, Y, N) getters.

    Y is (R, F) with R = N * n_blocks.  Returns a dict with
    S1 (F, p), ST (F, p), S2 (F, p, p), S1_agg (p,), ST_agg (p,),
    V1 (F, p), VT (F, p), var (F,)  -- all float64, accumulated in longdouble.
    """
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    R, F = Y.shape
    nb = R // N
    assert nb * N == R and nb >= p + 2
    Yc = Y.astype(LD) - Y.astype(LD).mean(axis=0)          # centring over all rows
    yA, yB = Yc[0:N], Yc[N:2 * N]
    var = yA.var(axis=0, ddof=1)
    f02 = _dot(yA, yB) / N
    muA = yA.mean(axis=0)
    V1 = np.zeros((F, p), dtype=LD)
    VT = np.zeros((F, p), dtype=LD)
    for i in range(p):
        yE = Yc[(2 + i) * N:(3 + i) * N]
        V1[:, i] = _dot(yB, yE) / (N - 1) - f02
        VT[:, i] = var - _dot(yA, yE) / (N - 1) + muA * muA
    S1 = V1 / var[:, None]
    ST = VT / var[:, None]
    S2 = np.zeros((F, p, p), dtype=LD)
    if second_order:
        for i in range(p):
            yE = Yc[(2 + i) * N:(3 + i) * N]
            for j in range(i):
                if p > 2:
                    yC = Yc[(2 + p + j) * N:(3 + p + j) * N]
                else:  # p == 2: C^j is E^(1-j)
                    yC = Yc[(2 + (1 - j)) * N:(3 + (1 - j)) * N]
                s = (_dot(yE, yC) / (N - 1) - f02) / var - S1[:, i] - S1[:, j]
                S2[:, i, j] = s
                S2[:, j, i] = s
    out = dict(S1=S1, ST=ST, S2=S2, V1=V1, VT=VT, var=var,
               S1_agg=V1.sum(axis=0) / var.sum(), ST_agg=VT.sum(axis=0) / var.sum())
    return {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}


def jansen_indices(Y, N, p):
    """Jansen 1999 estimators on the same design (OT JansenSensitivityAlgorithm):
    V_i = Var - sum((yB - yE^i)^2)/(2N-1);  VT_i = sum((yA - yE^i)^2)/(2N-1)."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    Yc = Y.astype(LD) - Y.astype(LD).mean(axis=0)
    yA, yB = Yc[0:N], Yc[N:2 * N]
    var = yA.var(axis=0, ddof=1)
    F = Y.shape[1]
    V1 = np.zeros((F, p), dtype=LD)
    VT = np.zeros((F, p), dtype=LD)
    for i in range(p):
        yE = Yc[(2 + i) * N:(3 + i) * N]
        V1[:, i] = var - ((yB - yE) ** 2).sum(axis=0) / (2 * N - 1)
        VT[:, i] = ((yA - yE) ** 2).sum(axis=0) / (2 * N - 1)
    out = dict(S1=V1 / var[:, None], ST=VT / var[:, None], V1=V1, VT=VT, var=var,
               S1_agg=V1.sum(axis=0) / var.sum(), ST_agg=VT.sum(axis=0) / var.sum())
    return {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}


def martinez_indices(Y, N, p):
    """Martinez 2011 estimators on the same design (OT MartinezSensitivityAlgorithm; the reference's comment at
    uq/uq.py:340 and doc/technical_documentation/uq.rst:64 name it):  S_i = corr(yB, yE^i),  ST_i = 1 - corr(yA, yE^i)
    (Pearson correlations over the N base rows: invariant to the centring and to the variance normalisation).
    Aggregation as for every OT estimator: sum_q S_i^q Var^q / sum_q Var^q.  PARITY UNPINNED (module header)."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    Yl = Y.astype(LD)
    yA, yB = Yl[0:N], Yl[N:2 * N]
    var = yA.var(axis=0, ddof=1)
    F = Y.shape[1]
    S1 = np.zeros((F, p), dtype=LD)
    ST = np.zeros((F, p), dtype=LD)
    dA, dB = yA - yA.mean(axis=0), yB - yB.mean(axis=0)
    for i in range(p):
        yE = Yl[(2 + i) * N:(3 + i) * N]
        dE = yE - yE.mean(axis=0)
        S1[:, i] = (dB * dE).sum(axis=0) / np.sqrt((dB * dB).sum(axis=0) * (dE * dE).sum(axis=0))
        ST[:, i] = 1 - (dA * dE).sum(axis=0) / np.sqrt((dA * dA).sum(axis=0) * (dE * dE).sum(axis=0))
    out = dict(S1=S1, ST=ST, var=var, S1_agg=(S1 * var[:, None]).sum(axis=0) / var.sum(),
               ST_agg=(ST * var[:, None]).sum(axis=0) / var.sum())
    return {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}


def mauntz_kucherenko_indices(Y, N, p):
    """Sobol'-Mauntz-Kucherenko 2007 estimators (OT MauntzKucherenkoSensitivityAlgorithm) on the centred output:
    V_i = sum(yB (yE^i - yA)) / (N - 1),  VT_i = sum(yA (yA - yE^i)) / (N - 1).  PARITY UNPINNED (module header)."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    Yc = Y.astype(LD) - Y.astype(LD).mean(axis=0)
    yA, yB = Yc[0:N], Yc[N:2 * N]
    var = yA.var(axis=0, ddof=1)
    F = Y.shape[1]
    V1 = np.zeros((F, p), dtype=LD)
    VT = np.zeros((F, p), dtype=LD)
    for i in range(p):
        yE = Yc[(2 + i) * N:(3 + i) * N]
        V1[:, i] = (yB * (yE - yA)).sum(axis=0) / (N - 1)
        VT[:, i] = (yA * (yA - yE)).sum(axis=0) / (N - 1)
    out = dict(S1=V1 / var[:, None], ST=VT / var[:, None], V1=V1, VT=VT, var=var,
               S1_agg=V1.sum(axis=0) / var.sum(), ST_agg=VT.sum(axis=0) / var.sum())
    return {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}


def asymptotic_sigma_aggregated(Y, N, p):
    """Standard deviations of the asymptotic (delta-method) distribution of the aggregated Saltelli indices: what
    ``ot.ResourceMap.SetAsBool('SobolIndicesAlgorithm-DefaultUseAsymptoticDistribution', True)`` selects for
    ``getFirstOrderIndicesInterval()`` / ``getTotalOrderIndicesInterval()`` (uq/uq.py:341, 425-426).

    Restated from the OpenTURNS 1.15 documentation of SaltelliSensitivityAlgorithm (PARITY UNPINNED, module header):
    with the centred output, per base row k and aggregated over the outputs q,
        first order  U_k = (sum_q yB_kq yE^i_kq, sum_q yA_kq^2),  psi(x, y) = x / y
        total order  U_k = (sum_q yA_kq yE^i_kq, sum_q yA_kq^2),  psi(x, y) = 1 - x / y
    and Var = grad psi(mean U)^T Cov(U) grad psi(mean U) / N with the unbiased sample covariance of U.
    Returns (sigma_S1 (p,), sigma_ST (p,)); the 95 % interval is index -+ 1.959964 sigma."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    Yc = Y.astype(LD) - Y.astype(LD).mean(axis=0)
    yA, yB = Yc[0:N], Yc[N:2 * N]
    b = (yA * yA).sum(axis=1)

    def sigma(a):
        x, y = a.mean(), b.mean()
        cxx = ((a - x) ** 2).sum() / (N - 1)
        cyy = ((b - y) ** 2).sum() / (N - 1)
        cxy = ((a - x) * (b - y)).sum() / (N - 1)
        return np.sqrt((cxx / y ** 2 - 2 * x * cxy / y ** 3 + x ** 2 * cyy / y ** 4) / N)
    s1 = np.zeros(p, dtype=LD)
    st = np.zeros(p, dtype=LD)
    for i in range(p):
        yE = Yc[(2 + i) * N:(3 + i) * N]
        s1[i] = sigma((yB * yE).sum(axis=1))
        st[i] = sigma((yA * yE).sum(axis=1))
    return np.asarray(s1, dtype=np.float64), np.asarray(st, dtype=np.float64)


def jackknife_aggregated(Y, N, p, groups, second_order=True):
    """Delete-a-group jackknife standard errors of the aggregated first / total order indices.

    NOT the reference's estimator: batman asks OpenTURNS for asymptotic (delta-method) intervals
    (uq/uq.py:341, 425-426), whose formulas cannot be pinned here (module header).  The drop-in reports
    intervals from a delete-a-group jackknife over disjoint groups of base rows -- asymptotically equivalent
    for a smooth function of sample means.  ``groups`` is an (N,) array of group ids; groups may differ in
    size (weights h_g = N / n_g, Busing, Meijer & van der Leeden 1999):

        theta_J  = G theta - sum_g (1 - 1/h_g) theta_(g)
        var_J    = 1/G sum_g (h_g theta - (h_g - 1) theta_(g) - theta_J)^2 / (h_g - 1)

    Returns (theta (2p,), sigma (2p,)) with theta = [S1_agg, ST_agg]."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    groups = np.asarray(groups)
    nb = Y.shape[0] // N
    full = saltelli_indices(Y, N, p, second_order)
    theta = np.concatenate([full['S1_agg'], full['ST_agg']])
    ids = np.unique(groups)
    G = len(ids)
    th_g, h = [], []
    for g in ids:
        keep = groups != g
        rows = np.concatenate([np.flatnonzero(keep) + b * N for b in range(nb)])
        est = saltelli_indices(Y[rows], int(keep.sum()), p, second_order)
        th_g.append(np.concatenate([est['S1_agg'], est['ST_agg']]))
        h.append(N / float((~keep).sum()))
    th_g, h = np.array(th_g), np.array(h)[:, None]
    theta_j = G * theta - ((1.0 - 1.0 / h) * th_g).sum(axis=0)
    pseudo = h * theta - (h - 1.0) * th_g
    var = (((pseudo - theta_j) ** 2) / (h - 1.0)).sum(axis=0) / G
    return theta, np.sqrt(var)


def uq_sobol_aggregate(Y, N, p, indices='aggregated', second_order=True):
    """Return value of ``UQ.sobol()``: [S2, S1, ST]  (uq/uq.py:400-436, 467-471, 499).

    * 'aggregated': S2 is the variance-weighted mean of the per-output S2 with
      ``output_var = output_design.var(axis=0)`` -- ddof=0 over ALL rows
      (uq.py:403, 412-420) after ``nan_to_num``; S1/ST are OT's aggregated
      getters (uq.py:423-424).
    * 'block': Y must already be the trapz-integrated scalar (uq.py:205-218);
      returns [S2[0], S1[0], ST[0]] (uq.py:466-468).
    """
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[:, None]
    est = saltelli_indices(Y, N, p, second_order)
    if indices == 'block':
        return [est['S2'][0], est['S1'][0], est['ST'][0]]
    output_var = Y.astype(LD).var(axis=0)                                  # uq.py:403
    s2 = np.zeros((p, p), dtype=LD)
    for q in range(Y.shape[1]):
        s2 += output_var[q] * np.nan_to_num(est['S2'][q])                  # uq.py:414-415
    s2 = s2 / output_var.sum()                                             # uq.py:418-420
    return [np.asarray(s2, dtype=np.float64), est['S1_agg'], est['ST_agg']]


def int_func(Y):
    """UQ.int_func: ``np.trapz(f_eval[0])`` with unit spacing (uq.py:216-217).

    numpy's trapz is ``(1.0 * (y[1:] + y[:-1]) / 2.0).sum()``; for a scalar
    output (F == 1) that is an empty sum, i.e. 0.0 -- the reference's docstring
    says "returns the prediction" but the code returns 0; restated as coded.
    """
    Y = np.asarray(Y, dtype=np.float64)
    return ((Y[:, 1:] + Y[:, :-1]) / 2.0).sum(axis=1)[:, None]
