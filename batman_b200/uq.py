"""``UQ`` drop-in: ``sobol()``, the propagation sample and its moments, ``error_model``
(uq/uq.py:79-185, 192-284, 286-529 of the reference).

The reference draws a Saltelli experiment with OpenTURNS, evaluates the
surrogate on its N(2p+2) rows one point at a time and hands the outputs to
``ot.SaltelliSensitivityAlgorithm``.  Here the three steps are one device pass:
base rows are generated in-kernel (or supplied), predicted, multiplied into the
estimator's sums and reduced, so neither the design nor the field outputs are
stored.  The per-block surrogate outputs (a few doubles per row) are kept in
HBM for the confidence intervals, which need the exact centring that is only
known after the pass.

Return value and conventions are the reference's: ``sobol()`` returns
``[S2, S1, ST]`` -- aggregated over the outputs for ``indices='aggregated'``
(uq.py:400-436), of the trapz-integrated output for ``'block'`` (uq.py:319-324,
466-468).  Estimator conventions: oracle/saltelli.py header (OpenTURNS itself
is not installable here: parity with OT's numbers is unpinned, parity with the
restated estimators is tested to 1e-8).

Base rows shard across ranks when ``torch.distributed`` is initialised: rank g
takes rows [g N/G, (g+1) N/G), the per-rank (hi, lo) sums are summed with ONE
all-reduce (a second one for the interval sums), every rank finalises.
"""
import ctypes
import logging
import os

import numpy as np

from . import _lib, distributions

ESTIMATORS = ('saltelli', 'jansen', 'martinez', 'mauntz-kucherenko')
_KEEP_LIMIT_BYTES = 16 << 30       # kept block outputs beyond this fall back to the jackknife intervals
_Z95 = 1.959963984540054           # 95 %, two-sided (OpenTURNS SobolIndicesAlgorithm-DefaultConfidenceLevel)


def shard_rows(size, world, rank):
    """Base rows of rank `rank`: contiguous slab [lo, hi); the slabs partition [0, size)."""
    return rank * size // world, (rank + 1) * size // world


def allreduce_totals(totals):
    """Sum the per-rank (hi, lo) Saltelli totals over all ranks -- the path's collective
    (NCCL over NVLink for CUDA tensors; gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    return totals


def asymptotic_sigma(sums, n, p):
    """Delta-method standard deviations of the aggregated first / total order indices from the row sums of
    ``bk_sobol_asymptotic`` ((2 + 6p,) values: b, b^2, then per input a1, a1^2, a1 b, a2, a2^2, a2 b): the variance of
    psi(x, y) = x / y (total order 1 - x / y) at the mean of U = (a, b), Cov(U) the unbiased sample covariance,
    divided by n (OpenTURNS SaltelliSensitivityAlgorithm asymptotic distribution; oracle/saltelli.py)."""
    sums = np.asarray(sums, dtype=np.float64)
    y = sums[0] / n
    cyy = (sums[1] - n * y * y) / (n - 1)
    out = np.empty((2, p))
    for i in range(p):
        for t in range(2):
            sa, saa, sab = sums[2 + 6 * i + 3 * t:2 + 6 * i + 3 * t + 3]
            x = sa / n
            cxx = (saa - n * x * x) / (n - 1)
            cxy = (sab - n * x * y) / (n - 1)
            out[t, i] = np.sqrt(max(0.0, (cxx / y ** 2 - 2 * x * cxy / y ** 3 + x ** 2 * cyy / y ** 4) / n))
    return out[0], out[1]


class UQ:
    """Uncertainty Quantification class (Sobol' indices, propagation moments, surrogate error)."""

    logger = logging.getLogger(__name__)

    def __init__(self, surrogate, dists=None, nsample=1000, method='sobol', indices='aggregated', space=None,
                 data=None, plabels=None, xlabel=None, flabel=None, xdata=None, fname=None, test=None, mesh={},
                 seed=123456, base_samples=None, second_order=True, shard=True, estimator='saltelli',
                 intervals='auto'):
        """Arguments as uq.py:79-119.  Additions, all keyword-only in spirit:

        :param bool second_order: ``True`` (what the reference always does,
          ``SobolIndicesExperiment(dist, N, True)``, uq.py:331): design with the
          C blocks, N(2p+2) predictions.  ``False``: first and total order only,
          N(p+2) predictions, the returned second-order matrix is zero.
        :param int seed: seed of the counter-based base-sample generator
          (batman seeds OpenTURNS with 123456, batman/__init__.py:6).
        :param base_samples: optional ``(A, B)`` arrays (nsample, p) to use instead
          of the generator (any distribution, any RNG).
        :param bool shard: with ``torch.distributed`` initialised, split the base
          rows over the ranks (default); ``False`` keeps the whole experiment on
          this process' GPU (no collective: what the sharded result is checked
          against).
        :param str estimator: which OpenTURNS estimator ``sobol()`` returns first
          and total order from: ``'saltelli'`` (the reference, uq.py:342),
          ``'jansen'``, ``'martinez'`` or ``'mauntz-kucherenko'`` (the comment at
          uq.py:340); all four come out of the same pass (``details``).
        :param str intervals: ``'asymptotic'`` -- the delta-method intervals the
          reference asks OpenTURNS for (uq.py:341, 425-426); ``'jackknife'`` -- a
          delete-a-group jackknife over the partial sums; ``'auto'`` (default)
          -- asymptotic unless the kept outputs would exceed 16 GB per GPU.
        """
        self.logger.info("\n----- UQ module -----")
        if estimator not in ESTIMATORS:
            raise ValueError("estimator must be one of %r" % (ESTIMATORS,))
        if intervals not in ('auto', 'asymptotic', 'jackknife'):
            raise ValueError("intervals must be 'auto', 'asymptotic' or 'jackknife'")
        self.test = test
        self.fname = fname
        self.xlabel = xlabel
        self.flabel = flabel
        self.mesh_kwargs = mesh
        self.surrogate = surrogate
        self.p_len = len(dists)
        self.plabels = ["x" + str(i) for i in range(self.p_len)] if plabels is None else plabels
        self.method_sobol = method
        self.type_indices = indices
        self.space = space
        self.points_sample = nsample
        self.seed = int(seed)
        self.base_samples = base_samples
        self.second_order = bool(second_order)
        self.shard = bool(shard)
        self.estimator = estimator
        self.intervals = intervals
        self.dists = dists
        try:
            self.marg_kinds, self.marg_params = distributions.parse_all(dists)
        except SystemError:
            self.logger.error('OpenTURNS distribution unknown.')
            raise
        except NotImplementedError:
            if surrogate is not None and base_samples is None:
                raise
            # explicit base samples / ensemble mode: the marginals are only needed for the propagation sample
            self.marg_kinds, self.marg_params = None, None
        if self.surrogate is not None:
            if self.surrogate.pod is not None:
                self.output_len = int(self.surrogate.pod.mean_snapshot.shape[0])
            else:
                self.output_len = int(self.surrogate.predictor.model_len)
            self.init_size = self.surrogate.space.doe_init
        else:                                                            # ensemble mode, uq.py:175-180
            self.init_size = len(space)
            data = np.asarray(data, dtype=np.float64)
            self.output_len = data.shape[1]
            self._output = data
            self.points_sample = self.init_size
        if (xdata is None) and (self.output_len > 1):
            self.xdata = np.linspace(0, 1, self.output_len)
        else:
            self.xdata = xdata
        self.details = None
        self._sample = None
        self._sample_dev = None
        self._output_dev = None
        if surrogate is not None:
            self._output = None

    # ------------------------------------------------------------------ LHS propagation sample (uq.py:159-173)
    def _lhs_device(self):
        """N-point randomised LHS of the joint distribution (``ot.LHSExperiment(dist, N, True, True)``,
        uq.py:159-160) on the device: the cell permutation of every dimension is drawn on the host (seeded numpy --
        OpenTURNS' own stream cannot be reproduced), shifts and inverse CDFs in ``bk_lhs_rows``."""
        import torch
        if self._sample_dev is None:
            if self.marg_kinds is None:
                raise NotImplementedError("the propagation sample needs marginals with a device inverse CDF")
            n, p = int(self.points_sample), self.p_len
            rng = np.random.RandomState(self.seed)
            perm = np.stack([rng.permutation(n) for _ in range(p)], axis=1).astype(np.int64)
            dev = torch.device('cuda', torch.cuda.current_device())
            perm_d = torch.from_numpy(np.ascontiguousarray(perm)).to(dev)
            out = torch.empty((n, p), dtype=torch.float64, device=dev)
            _lib.check(_lib.load().bk_lhs_rows(self.seed, n, p, _lib.as_int_array(self.marg_kinds),
                                               _lib.as_double_array(self.marg_params), ctypes.c_void_p(perm_d.data_ptr()),
                                               ctypes.c_void_p(out.data_ptr()),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            self._sample_dev, self._lhs_perm = out, perm
        return self._sample_dev

    @property
    def sample(self):
        """The propagation sample as a host array (uq.py:159); in ensemble mode the user's space (uq.py:175)."""
        if self.surrogate is None:
            return self.space
        if self._sample is None:
            self._sample = self._lhs_device().cpu().numpy()
        return self._sample

    def _output_device(self):
        """Surrogate mean on the propagation sample (uq.py:173), resident on the device."""
        if self._output_dev is None:
            self._output_dev, _ = self.surrogate.predict_device(self._lhs_device(), return_std=False)
        return self._output_dev

    @property
    def output(self):
        if self._output is None and self.surrogate is not None:
            self._output = self._output_device().cpu().numpy()
        return self._output

    @output.setter
    def output(self, value):
        self._output = value

    def moments(self):
        """Mean, standard deviation (ddof = 0), minimum and maximum of every output over the propagation sample: what
        ``error_propagation`` hands to the PDF plot (uq.py:501-529 -> visualization/uncertainty.py:136-143).  The
        sample is predicted and reduced on the device (``bk_moments``: compensated sums around the first row); only
        4 x output_len numbers come back."""
        import torch
        if self.surrogate is None:
            y = torch.from_numpy(np.ascontiguousarray(self._output, dtype=np.float64)).cuda()
        else:
            y = self._output_device()
        y = y.contiguous()
        rows, F = y.shape
        shift = y[0].contiguous()                       # any pilot value close to the data keeps the squares small
        out = torch.empty((F, 6), dtype=torch.float64, device=y.device)
        with torch.cuda.device(y.device):
            _lib.check(_lib.load().bk_moments(ctypes.c_void_p(y.data_ptr()), rows, F, ctypes.c_void_p(shift.data_ptr()),
                                              ctypes.c_void_p(out.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        o = out.cpu().numpy().astype(np.longdouble)
        s, q = o[:, 0] + o[:, 1], o[:, 2] + o[:, 3]
        mean_c = s / rows
        var = np.maximum(q / rows - mean_c * mean_c, 0.0)
        mean = np.asarray(mean_c + shift.cpu().numpy().astype(np.longdouble), dtype=np.float64)
        self._moments = dict(mean=mean, sd=np.asarray(np.sqrt(var), dtype=np.float64),
                             min=np.asarray(o[:, 4], dtype=np.float64), max=np.asarray(o[:, 5], dtype=np.float64))
        return self._moments['mean'], self._moments['sd']

    def error_propagation(self):
        """Moments of the propagation sample (uq.py:501-529).  The reference goes on to plot PDFs and
        correlation matrices (visualization/, out of scope); the numbers it plots are returned and, with an output
        folder, written as ``moment.json``: {min, sd_min, mean, sd_max, max} per output (uncertainty.py:199-206)."""
        self.logger.info("\n----- Uncertainty Propagation -----")
        mean, sd = self.moments()
        res = dict(min=self._moments['min'], sd_min=mean - sd, mean=mean, sd_max=mean + sd, max=self._moments['max'])
        if self.fname is not None:
            import json
            os.makedirs(self.fname, exist_ok=True)
            with open(os.path.join(self.fname, 'moment.json'), 'w') as fd:
                json.dump({k: np.atleast_1d(v)[:, None].tolist() for k, v in res.items()}, fd)
        return res

    def __repr__(self):
        return ("UQ object: Method({}), Input({}), Distribution({})"
                .format(self.method_sobol, self.plabels, self.dists))

    # ------------------------------------------------------------------ surrogate evaluation
    def func(self, coords):
        """Surrogate mean at raw points: (n, output_len) (uq.py:192-203)."""
        f_eval, _ = self.surrogate(np.atleast_2d(coords))
        return f_eval

    def int_func(self, coords):
        """Trapezoidal integral of the surrogate mean along the output (uq.py:205-218)."""
        f_eval = self.func(coords)
        return ((f_eval[:, 1:] + f_eval[:, :-1]) / 2.0).sum(axis=1)[:, None]

    def error_model(self, indices, function):
        r"""Error between the surrogate and an analytic function (uq.py:220-284): Q2 and MSE on the propagation
        sample, L2 distances of the Sobol' indices to the closed-form ones; ``model_err.dat`` as the reference writes
        it.  ``function``: a name (``Ishigami``, ``G_Function``; any name of ``batman.functions.analytical`` when
        batman is importable) or an object with ``__call__``, ``s_first``, ``s_second``, ``s_total``.

        :return: err_q2, mse, s_l2_2nd, s_l2_1st, s_l2_total."""
        from sklearn.metrics import mean_squared_error, r2_score
        from . import functions
        fun = functions.resolve(function)
        s_l2_2nd = np.sqrt(np.sum((fun.s_second - indices[0]) ** 2))
        s_l2_1st = np.sqrt(np.sum((fun.s_first - indices[1]) ** 2))
        s_l2_total = np.sqrt(np.sum((fun.s_total - indices[2]) ** 2))
        y_ref = np.array(fun(self.sample))
        y_pred = np.array(self.output)                                  # self.func(self.sample), evaluated once
        err_q2 = r2_score(y_ref, y_pred, multioutput='uniform_average')
        mse = mean_squared_error(y_ref, y_pred, multioutput='uniform_average')
        self.logger.info("\n----- Surrogate Model Error -----\nQ2: {}\nMSE: {}\n"
                         "L2(sobol 2nd, 1st and total order indices error): {}, {}, {}"
                         .format(err_q2, mse, s_l2_2nd, s_l2_1st, s_l2_total))
        if self.fname is not None:
            os.makedirs(self.fname, exist_ok=True)
            with open(os.path.join(self.fname, 'model_err.dat'), 'w') as f:
                f.writelines("{} {} {} {} {} {}".format(self.init_size, err_q2, self.points_sample, s_l2_2nd, s_l2_1st,
                                                        s_l2_total))
        else:
            self.logger.debug("No output folder to write errors in")
        self.model_error = (err_q2, mse, s_l2_2nd, s_l2_1st, s_l2_total)
        return self.model_error

    # ------------------------------------------------------------------ Sobol'
    def sobol(self):
        """Compute Sobol' indices: returns [second, first, total] (uq.py:286-499)."""
        import torch
        import torch.distributed as dist
        if self.method_sobol != 'sobol':
            raise NotImplementedError("method %r: only the Saltelli 'sobol' path is accelerated" % self.method_sobol)
        if not torch.cuda.is_available():
            raise RuntimeError("batman_b200.UQ.sobol needs a CUDA device: there is no CPU fallback")
        self.logger.info("\n----- Sobol' indices -----")
        lib = _lib.load()
        p = self.p_len
        block = self.type_indices == 'block'
        F = 1 if block else self.output_len
        second = 1 if self.second_order else 0
        nsums = lib.bk_sobol_nsums(p, second)
        nslots = lib.bk_sobol_nslots()
        nb = 2 * p + 2 if (p > 2 and second) else p + 2
        dev = torch.device('cuda', torch.cuda.current_device())
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        partials = torch.zeros((nslots, F, nsums, 2), dtype=torch.float64, device=dev)
        world, rank = 1, 0
        if self.shard and self.surrogate is not None and dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()

        def ptr(t):
            return ctypes.c_void_p(t.data_ptr()) if t is not None else None

        # what the interval pass needs afterwards: [(stored outputs (nb, cnt, r), cnt)], the basis that lifts them to the
        # outputs the estimator saw (None = identity) and the pilot shift per stored output
        kept, keep_basis = [], None
        if self.surrogate is not None:
            size = int(self.points_sample)
            lo, hi = shard_rows(size, world, rank)
            if self.base_samples is None and self.marg_kinds is None:
                raise NotImplementedError("these marginals have no device inverse CDF: pass base_samples=(A, B)")
            kinds = _lib.as_int_array(self.marg_kinds) if self.marg_kinds is not None else None
            params = _lib.as_double_array(self.marg_params) if self.marg_kinds is not None else None
            A_all = B_all = None
            if self.base_samples is not None:
                A_all = torch.from_numpy(np.ascontiguousarray(self.base_samples[0], dtype=np.float64)).to(dev)
                B_all = torch.from_numpy(np.ascontiguousarray(self.base_samples[1], dtype=np.float64)).to(dev)
                if A_all.shape != (size, p) or B_all.shape != (size, p):
                    raise ValueError("base_samples must be two (nsample, p) arrays")
            pod = self.surrogate.pod
            scaler = self.surrogate.scaler
            m = int(self.surrogate.predictor.model_len)
            want_asym = self.type_indices == 'aggregated' or block
            keep = want_asym and self.intervals != 'jackknife' and \
                (self.intervals == 'asymptotic' or nb * (hi - lo) * m * 8 <= _KEEP_LIMIT_BYTES)
            # raw design rows go in: the MinMax scaling is fused for this call only (surrogate_model.py:181)
            with self.surrogate.predictor.input_scaler(scaler.scale_, scaler.min_) as model:
                if pod is None and not block:
                    # fused path: design -> mean -> products -> (hi, lo) sums; the field of outputs is never stored
                    shift = _lib.as_double_array(model.y_means)
                    ymeans_dev = torch.tensor(list(model.y_means), dtype=torch.float64, device=dev)   # before the long kernel
                    a_p = ptr(A_all[lo:hi]) if A_all is not None else None
                    b_p = ptr(B_all[lo:hi]) if B_all is not None else None
                    if keep and hi > lo:
                        ykeep = torch.empty((nb, hi - lo, m), dtype=torch.float64, device=dev)
                        _lib.check(lib.bk_sobol_accumulate_keep(model.handle, a_p, b_p, self.seed, kinds, params, lo,
                                                                hi - lo, second, shift, ptr(partials), ptr(ykeep), stream))
                        kept.append((ykeep, hi - lo))
                    else:
                        _lib.check(lib.bk_sobol_accumulate(model.handle, a_p, b_p, self.seed, kinds, params, lo, hi - lo,
                                                           second, shift, ptr(partials), stream))
                else:
                    # POD field / block integral: the per-block mode outputs of a chunk of base rows are stored
                    # (bk_sobol_modes), then lifted to the field on the tensor cores and reduced in one kernel
                    # (bk_pod_sobol_accumulate); the N(2p+2) x F output design never exists.
                    if pod is not None:
                        U_h, mean_h = pod.U, pod.mean_snapshot
                    else:
                        U_h, mean_h = np.eye(m), np.zeros(m)
                    if block:
                        # np.trapz(f) = sum((f[1:] + f[:-1]) / 2) is linear: integrate the basis once (uq.py:216-217)
                        w = np.zeros(len(mean_h))
                        if len(mean_h) > 1:
                            w[:-1] += 0.5
                            w[1:] += 0.5
                        U_h, mean_h = (w @ U_h)[None, :], np.array([w @ mean_h])
                    keep_basis = (np.asarray(U_h, dtype=np.float64), np.asarray(mean_h, dtype=np.float64))
                    shift_h = mean_h + U_h @ np.asarray(model.y_means)
                    U_d = torch.from_numpy(np.ascontiguousarray(U_h, dtype=np.float64)).to(dev)
                    mean_d = torch.from_numpy(np.ascontiguousarray(mean_h, dtype=np.float64)).to(dev)
                    shift_d = torch.from_numpy(np.ascontiguousarray(shift_h, dtype=np.float64)).to(dev)
                    chunk = max(1, min(hi - lo, max(1024, int((1 << 26) // max(1, nb * m)))))
                    ymodes = None
                    for c0 in range(lo, hi, chunk):
                        cnt = min(chunk, hi - c0)
                        if keep or ymodes is None:
                            ymodes = torch.empty(nb * cnt * m if keep else nb * chunk * m, dtype=torch.float64, device=dev)
                        _lib.check(lib.bk_sobol_modes(
                            model.handle, ptr(A_all[c0:c0 + cnt]) if A_all is not None else None,
                            ptr(B_all[c0:c0 + cnt]) if B_all is not None else None, self.seed, kinds, params, c0, cnt,
                            second, ptr(ymodes), stream))
                        _lib.check(lib.bk_pod_sobol_accumulate(ptr(ymodes), cnt, p, m, second, ptr(U_d), ptr(mean_d), F,
                                                               ptr(shift_d), ptr(partials), stream))
                        if keep:
                            kept.append((ymodes.view(nb, cnt, m), cnt))
            self.logger.info("Created {} samples for Sobol'".format(nb * size))
        else:
            # ensemble mode (uq.py:336-339): user-supplied Saltelli output sample.  For p == 2 the layout is the
            # N(p+2) rows OpenTURNS' experiment has (no C blocks); the reference divides by 2p+2 regardless.
            size = len(self.space) // nb
            Y = torch.from_numpy(np.ascontiguousarray(self._output[:nb * size], dtype=np.float64)).to(dev)
            if block:
                Y = ((Y[:, 1:] + Y[:, :-1]) / 2.0).sum(dim=1, keepdim=True).contiguous()
            shift_f = Y[:size].mean(dim=0).contiguous()
            _lib.check(lib.bk_sobol_accumulate_y(ptr(Y), size, size, p, F, second, ptr(shift_f), ptr(partials),
                                                 stream))
            if self.intervals != 'jackknife':
                kept.append((Y.view(nb, size, F), size))

        totals = torch.empty((F, nsums, 2), dtype=torch.float64, device=dev)
        _lib.check(lib.bk_sobol_reduce_slots(ptr(partials), F, nsums, ptr(totals), stream))
        if world > 1:
            allreduce_totals(totals)
        out = torch.empty(lib.bk_sobol_result_len(p, F), dtype=torch.float64, device=dev)
        _lib.check(lib.bk_sobol_finalize(ptr(totals), size, p, F, second, ptr(out), stream))
        # When the kept outputs ARE the estimator's outputs (fused path, ensemble mode) their centre is the mean over all
        # rows, which finalize just left on the device: the interval kernels and collectives are queued first and the
        # call ends with ONE device -> host read.  With a basis in between (POD field, block integral) the centre
        # needs a small host solve, so the indices are read first.
        nres = out.numel()
        center_dev = None
        if kept and keep_basis is None:
            center_dev = out[nres - F:] + (shift_f if self.surrogate is None else ymeans_dev)
        isums = None
        if center_dev is not None or not kept:
            abuf, jk = self._interval_sums(lib, partials, totals, kept, center_dev, None, size, p, F, second, nslots,
                                           world, stream)
            tail = [center_dev] if center_dev is not None else []
            host = torch.cat([out, abuf, jk.reshape(-1)] + tail).cpu().numpy()
            res, isums = host[:nres], host[nres:nres + abuf.numel() + jk.numel()]
        else:
            res = out.cpu().numpy()
        self.details = d = self._unpack(res, p, F)
        # mean of every estimator output over all rows (the centring of the estimator)
        if center_dev is not None:
            d['mean'] = host[nres + len(isums):].copy()
        elif self.surrogate is None:
            d['mean'] = d['mean'] + shift_f.cpu().numpy()
        elif keep_basis is None:
            d['mean'] = d['mean'] + np.asarray(self.surrogate.predictor._model().y_means)
        else:
            d['mean'] = d['mean'] + shift_h
        key = {'saltelli': ('S1', 'ST'), 'jansen': ('S1_jansen', 'ST_jansen'), 'martinez': ('S1_martinez', 'ST_martinez'),
               'mauntz-kucherenko': ('S1_mauntz_kucherenko', 'ST_mauntz_kucherenko')}[self.estimator]
        d['S1_selected'], d['ST_selected'] = d[key[0]], d[key[1]]
        d['S1_agg_selected'], d['ST_agg_selected'] = d[key[0] + '_agg'], d[key[1] + '_agg']
        if isums is None:
            center, R = self._interval_geometry(d, keep_basis, kept[0][0].shape[2])
            c_d = torch.from_numpy(np.ascontiguousarray(center)).to(dev)
            abuf, jk = self._interval_sums(lib, partials, totals, kept, c_d, R, size, p, F, second, nslots, world, stream)
            isums = torch.cat([abuf, jk.reshape(-1)]).cpu().numpy()
        self._interval_finish(isums, size, p, world)
        # the messages of uq.py:376-378, 427-436; formatted only when somebody listens (array formatting costs
        # milliseconds, more than a small Saltelli pass)
        if self.logger.isEnabledFor(logging.DEBUG):
            self.logger.debug("-> Second order:\n{}\n".format(d['S2']))
            self.logger.debug("-> First order:\n{}\n-> Total:\n{}\n".format(d['S1_selected'], d['ST_selected']))
        if self.type_indices == 'aggregated':
            aggregated = [d['S2_agg'], d['S1_agg_selected'], d['ST_agg_selected']]
            if self.logger.isEnabledFor(logging.INFO):
                self.logger.info("-> First order confidence:\n{}\n-> Total order confidence:\n{}\n"
                                 .format(d['S1_interval'], d['ST_interval']))
                self.logger.info("Aggregated_indices:\n-> Second order:\n{}\n-> First order:\n{}\n-> Total order:\n{}\n"
                                 .format(*aggregated))
        else:
            aggregated = [d['S2'][0], d['S1_selected'][0], d['ST_selected'][0]]            # uq.py:466-468
        self._write_files(d, p, F)
        if self.type_indices != 'aggregated':
            self.xdata = None
        # Compute error of the surrogate with a known function (uq.py:494-497)
        if self.type_indices in ('aggregated', 'block') and self.test and self.surrogate is not None:
            self.error_model(aggregated, self.test)
        return aggregated

    def _interval_sums(self, lib, partials, totals, kept, c_d, R, size, p, F, second, nslots, world, stream):
        """Device half of the 95 % intervals of the aggregated first / total order indices (uq.py:341, 425-426):
        -> (abuf, jk) device tensors, summed / gathered over the ranks, nothing read back.

        ``asymptotic``: the delta-method intervals the reference asks OpenTURNS for, from the kept block outputs and
        the exact centring ``c_d`` (``bk_sobol_asymptotic``; formulas restated in oracle/saltelli.py, unpinned like the
        rest of the OpenTURNS half).  ``jackknife``: delete-a-group jackknife over the partial-sum slots (disjoint
        groups of base rows, every rank's slots together), always computed as a cross-check (costs 0.5 ms at
        N = 1e6) and the fallback when the outputs were not kept.  ``abuf`` = [asymptotic sums | number of ranks that
        kept their outputs]: ONE all-reduce beside the all-gather of the jackknife groups."""
        import torch
        import torch.distributed as dist
        dev = partials.device
        jk = torch.empty((nslots, 2 + 2 * p), dtype=torch.float64, device=dev)
        _lib.check(lib.bk_sobol_jackknife(ctypes.c_void_p(partials.data_ptr()), ctypes.c_void_p(totals.data_ptr()),
                                          size, p, F, second, ctypes.c_void_p(jk.data_ptr()), stream))
        na = (2 + 6 * p) * 2
        abuf = torch.zeros(na + 1, dtype=torch.float64, device=dev)
        if kept:
            # asymptotic sums on this rank's kept outputs, accumulated in place at the head of abuf
            R_d = torch.from_numpy(np.ascontiguousarray(R)).to(dev) if R is not None else None
            zero_d = torch.zeros(R.shape[0] if R is not None else 1, dtype=torch.float64, device=dev)
            for z, cnt in kept:
                nbk, _, r = z.shape
                if R_d is not None:        # lift the stored modes to coordinates in which the outputs' dot product is plain
                    from .pod_kernels import pod_inverse
                    z = pod_inverse(zero_d, R_d, z.reshape(nbk * cnt, r)).view(nbk, cnt, R.shape[0])
                    r = R.shape[0]
                _lib.check(lib.bk_sobol_asymptotic(ctypes.c_void_p(z.data_ptr()), cnt, cnt, p, r,
                                                   ctypes.c_void_p(c_d.data_ptr()), ctypes.c_void_p(abuf.data_ptr()), stream))
            abuf[na] = 1.0
        if world > 1:
            parts = [torch.empty_like(jk) for _ in range(world)]
            dist.all_gather(parts, jk)
            jk = torch.cat(parts, dim=0)
            dist.all_reduce(abuf, op=dist.ReduceOp.SUM)
        return abuf, jk

    def _interval_finish(self, host, size, p, world):
        """Host half: interval widths from the summed interval sums (``host`` = abuf followed by the jackknife rows)."""
        d = self.details
        na = (2 + 6 * p) * 2
        asym = host[:na].reshape(2 + 6 * p, 2) if host[na] == float(world) else None   # every rank kept its outputs, or none is used
        jk = host[na + 1:].reshape(-1, 2 + 2 * p)
        jk = jk[jk[:, 0] > 0]
        theta = np.concatenate([d['S1_agg'], d['ST_agg']])
        G = len(jk)
        if G < 2:
            sigma_jk = np.full(2 * p, np.nan)
        else:
            h = (size / jk[:, 1])[:, None]
            th_g = jk[:, 2:]
            theta_j = G * theta - ((1.0 - 1.0 / h) * th_g).sum(axis=0)
            pseudo = h * theta - (h - 1.0) * th_g
            sigma_jk = np.sqrt((((pseudo - theta_j) ** 2) / (h - 1.0)).sum(axis=0) / G)
        d['S1_agg_sigma_jackknife'], d['ST_agg_sigma_jackknife'] = sigma_jk[:p], sigma_jk[p:]
        d['jackknife_groups'] = G
        if asym is not None:
            s1, st = asymptotic_sigma(asym[:, 0] + asym[:, 1], size, p)
            d['S1_agg_sigma_asymptotic'], d['ST_agg_sigma_asymptotic'] = s1, st
            d['interval_method'] = 'asymptotic'
            sigma = np.concatenate([s1, st])
        else:
            d['interval_method'] = 'jackknife'
            sigma = sigma_jk
        d['S1_agg_sigma'], d['ST_agg_sigma'] = sigma[:p], sigma[p:]
        c1, ct = d['S1_agg_selected'], d['ST_agg_selected']
        d['S1_interval'] = (c1 - _Z95 * sigma[:p], c1 + _Z95 * sigma[:p])
        d['ST_interval'] = (ct - _Z95 * sigma[p:], ct + _Z95 * sigma[p:])

    @staticmethod
    def _interval_geometry(d, keep_basis, r):
        """-> (centre (r',), R (r', r) or None): the kept outputs z (r per row) enter the interval sums as R z - centre,
        such that the plain dot product of two transformed rows equals the dot product over the estimator's outputs.
        No basis: the kept outputs ARE the estimator's outputs, centre = their mean over all rows.  With a basis
        y = mean_h + U_h z (POD field, or its trapz integral): <y1 - mu, y2 - mu> = (z1 - zbar)^T G (z2 - zbar) with
        G = U_h^T U_h = R^T R and mu = mean_h + U_h zbar."""
        if keep_basis is None:
            return np.asarray(d['mean'], dtype=np.float64), None
        U_h, mean_h = keep_basis
        if U_h.shape[0] == 1:                                  # block integral: one scalar per row
            return np.asarray(d['mean'] - mean_h, dtype=np.float64), U_h
        G = U_h.T @ U_h
        zbar = np.linalg.lstsq(U_h, d['mean'] - mean_h, rcond=None)[0]
        if np.allclose(G, np.eye(r), rtol=0, atol=1e-12):      # orthonormal POD basis: mode space is already Euclidean
            return zbar, None
        lam, V = np.linalg.eigh(G)
        R = (np.sqrt(np.maximum(lam, 0.0))[:, None] * V.T)
        return R @ zbar, R

    def _write_files(self, d, p, F):
        """sensitivity.json / sensitivity_aggregated.json as the reference writes them (uq.py:380-393, 437-463;
        layout of input_output/formater.py:61-77: {name: [[value], ...]})."""
        if self.fname is None:
            self.logger.debug("No output folder to write indices in")
            return
        import json

        def write(path, data, names):
            data = np.atleast_2d(data)
            with open(path, 'w') as fd:
                json.dump({name: data[:, j:j + 1].tolist() for j, name in enumerate(names)}, fd)
        os.makedirs(self.fname, exist_ok=True)
        block = self.type_indices == 'block'
        data = np.append(np.reshape(d['S1_selected'], (F, p)), np.reshape(d['ST_selected'], (F, p)), axis=1)
        names = ['S_{}'.format(lab) for lab in self.plabels] + ['S_T_{}'.format(lab) for lab in self.plabels]
        if (self.output_len > 1) and not block:
            names = ['x'] + names
            data = np.append(np.reshape(self.xdata, (F, 1)), data, axis=1)
        write(os.path.join(self.fname, 'sensitivity.json'), data, names)
        if self.type_indices == 'aggregated':
            (i1_min, i1_max), (i2_min, i2_max) = d['S1_interval'], d['ST_interval']
            agg = np.array([i1_min, d['S1_agg_selected'], i1_max, i2_min, d['ST_agg_selected'], i2_max]).flatten()
            names = [a + str(lab) for a in ['S_min_', 'S_', 'S_max_', 'S_T_min_', 'S_T_', 'S_T_max_'] for lab in self.plabels]
            write(os.path.join(self.fname, 'sensitivity_aggregated.json'), agg, names)

    @staticmethod
    def _design_blocks(A, B, nb):
        """[A; B; E^1..E^p; (C^1..C^p)] as a (nb, cnt, p) tensor (SobolIndicesExperiment layout)."""
        import torch
        cnt, p = A.shape
        D = torch.empty((nb, cnt, p), dtype=torch.float64, device=A.device)
        D[0], D[1] = A, B
        for i in range(p):
            D[2 + i] = A
            D[2 + i][:, i] = B[:, i]
        if nb == 2 * p + 2:
            for i in range(p):
                D[2 + p + i] = B
                D[2 + p + i][:, i] = A[:, i]
        return D

    @staticmethod
    def _unpack(res, p, F):
        """Packed result of ``bk_sobol_finalize`` (layout: include/batman_b200.h)."""
        o = 0

        def take(shape):
            nonlocal o
            n = int(np.prod(shape))
            v = res[o:o + n].reshape(shape)
            o += n
            return v
        d = {}
        d['S1'], d['ST'], d['S2'] = take((F, p)), take((F, p)), take((F, p, p))
        d['V1'], d['VT'], d['var'], d['output_var'] = take((F, p)), take((F, p)), take((F,)), take((F,))
        d['S1_jansen'], d['ST_jansen'] = take((F, p)), take((F, p))
        d['S1_agg'], d['ST_agg'], d['S2_agg'] = take((p,)), take((p,)), take((p, p))
        d['S1_martinez'], d['ST_martinez'] = take((F, p)), take((F, p))
        d['S1_mauntz_kucherenko'], d['ST_mauntz_kucherenko'] = take((F, p)), take((F, p))
        d['S1_jansen_agg'], d['ST_jansen_agg'] = take((p,)), take((p,))
        d['S1_martinez_agg'], d['ST_martinez_agg'] = take((p,)), take((p,))
        d['S1_mauntz_kucherenko_agg'], d['ST_mauntz_kucherenko_agg'] = take((p,)), take((p,))
        d['mean'] = take((F,)).copy()
        return d
