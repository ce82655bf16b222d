"""``UQ.sobol()`` drop-in (uq/uq.py:79-185, 192-218, 286-499 of the reference).

The reference draws a Saltelli experiment with OpenTURNS, evaluates the
surrogate on its N(2p+2) rows one point at a time and hands the outputs to
``ot.SaltelliSensitivityAlgorithm``.  Here the three steps are one device pass:
base rows are generated in-kernel (or supplied), predicted, multiplied into the
estimator's sums and reduced, so neither the design nor the outputs are stored.

Return value and conventions are the reference's: ``sobol()`` returns
``[S2, S1, ST]`` -- aggregated over the outputs for ``indices='aggregated'``
(uq.py:400-436), of the trapz-integrated output for ``'block'`` (uq.py:319-324,
466-468).  Estimator conventions: oracle/saltelli.py header (OpenTURNS itself
is not installable here: parity with OT's numbers is unpinned, parity with the
restated estimator is tested to 1e-8).

Base rows shard across ranks when ``torch.distributed`` is initialised: rank g
takes rows [g N/G, (g+1) N/G), the per-rank (hi, lo) sums are summed with ONE
all-reduce, every rank finalises.
"""
import ctypes
import logging

import numpy as np

from . import _lib, distributions


def shard_rows(size, world, rank):
    """Base rows of rank `rank`: contiguous slab [lo, hi); the slabs partition [0, size)."""
    return rank * size // world, (rank + 1) * size // world


def allreduce_totals(totals):
    """Sum the per-rank (hi, lo) Saltelli totals over all ranks -- the path's only collective
    (NCCL over NVLink for CUDA tensors; gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(totals, op=dist.ReduceOp.SUM)
    return totals


class UQ:
    """Uncertainty Quantification class (Sobol' part)."""

    logger = logging.getLogger(__name__)

    def __init__(self, surrogate, dists=None, nsample=1000, method='sobol', indices='aggregated', space=None,
                 data=None, plabels=None, xlabel=None, flabel=None, xdata=None, fname=None, test=None, mesh={},
                 seed=123456, base_samples=None, second_order=True, shard=True):
        """Arguments as uq.py:79-119.  Three additions, all keyword-only in spirit:

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
        """
        self.logger.info("\n----- UQ module -----")
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
        self.dists = dists
        if base_samples is None or surrogate is None:
            try:
                self.marg_kinds, self.marg_params = distributions.parse_all(dists)
            except SystemError:
                self.logger.error('OpenTURNS distribution unknown.')
                raise
        else:
            self.marg_kinds, self.marg_params = [0] * self.p_len, [0.0, 1.0] * self.p_len
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
        if surrogate is not None:
            self._output = None

    # ------------------------------------------------------------------ LHS propagation sample (uq.py:159-173)
    @property
    def sample(self):
        """N-point randomised LHS of the joint distribution (``ot.LHSExperiment(dist, N, True, True)``,
        uq.py:159-160), drawn lazily with numpy (seeded) -- OpenTURNS' stream cannot be reproduced."""
        if self._sample is None and self.surrogate is not None:
            from scipy.special import ndtri
            rng = np.random.RandomState(self.seed)
            n, p = int(self.points_sample), self.p_len
            u = np.empty((n, p))
            for j in range(p):
                u[:, j] = (rng.permutation(n) + rng.rand(n)) / n        # shuffled cells, random shift
            x = np.empty_like(u)
            for j in range(p):
                a, b = self.marg_params[2 * j], self.marg_params[2 * j + 1]
                x[:, j] = a + (b - a) * u[:, j] if self.marg_kinds[j] == distributions.UNIFORM else a + b * ndtri(u[:, j])
            self._sample = x
        return self._sample if self.surrogate is not None else self.space

    @property
    def output(self):
        """Surrogate mean on the LHS sample (uq.py:173), evaluated on the device on first use."""
        if self._output is None and self.surrogate is not None:
            self._output = self.func(self.sample)
        return self._output

    @output.setter
    def output(self, value):
        self._output = value

    def moments(self):
        """Mean and standard deviation of every output over the LHS sample (what error_propagation
        plots, uq.py:501-529)."""
        out = np.asarray(self.output, dtype=np.float64)
        return out.mean(axis=0), out.std(axis=0)

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

        if self.surrogate is not None:
            size = int(self.points_sample)
            lo, hi = shard_rows(size, world, rank)
            kinds = _lib.as_int_array(self.marg_kinds)
            params = _lib.as_double_array(self.marg_params)
            A_all = B_all = None
            if self.base_samples is not None:
                A_all = torch.from_numpy(np.ascontiguousarray(self.base_samples[0], dtype=np.float64)).to(dev)
                B_all = torch.from_numpy(np.ascontiguousarray(self.base_samples[1], dtype=np.float64)).to(dev)
                if A_all.shape != (size, p) or B_all.shape != (size, p):
                    raise ValueError("base_samples must be two (nsample, p) arrays")
            pod = self.surrogate.pod
            scaler = self.surrogate.scaler
            # raw design rows go in: the MinMax scaling is fused for this call only (surrogate_model.py:181)
            with self.surrogate.predictor.input_scaler(scaler.scale_, scaler.min_) as model:
                if pod is None and not block:
                    # fused path: design -> mean -> products -> (hi, lo) sums, nothing stored
                    shift = _lib.as_double_array(model.y_means)
                    _lib.check(lib.bk_sobol_accumulate(
                        model.handle, ptr(A_all[lo:hi]) if A_all is not None else None,
                        ptr(B_all[lo:hi]) if B_all is not None else None, self.seed, kinds, params, lo, hi - lo,
                        second, shift, ptr(partials), stream))
                else:
                    # POD field / block integral: the per-block mode outputs of a chunk of base rows are stored
                    # (bk_sobol_modes), then lifted to the field on the tensor cores and reduced in one kernel
                    # (bk_pod_sobol_accumulate); the N(2p+2) x F output design never exists.
                    m = model.m
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
                    shift_h = mean_h + U_h @ np.asarray(model.y_means)
                    U_d = torch.from_numpy(np.ascontiguousarray(U_h, dtype=np.float64)).to(dev)
                    mean_d = torch.from_numpy(np.ascontiguousarray(mean_h, dtype=np.float64)).to(dev)
                    shift_d = torch.from_numpy(np.ascontiguousarray(shift_h, dtype=np.float64)).to(dev)
                    chunk = max(1, min(hi - lo, max(1024, int((1 << 26) // max(1, nb * m)))))
                    ymodes = torch.empty(nb * chunk * m, dtype=torch.float64, device=dev)
                    for c0 in range(lo, hi, chunk):
                        cnt = min(chunk, hi - c0)
                        _lib.check(lib.bk_sobol_modes(
                            model.handle, ptr(A_all[c0:c0 + cnt]) if A_all is not None else None,
                            ptr(B_all[c0:c0 + cnt]) if B_all is not None else None, self.seed, kinds, params, c0, cnt,
                            second, ptr(ymodes), stream))
                        _lib.check(lib.bk_pod_sobol_accumulate(ptr(ymodes), cnt, p, m, second, ptr(U_d), ptr(mean_d), F,
                                                               ptr(shift_d), ptr(partials), stream))
            self.logger.info("Created {} samples for Sobol'".format(nb * size))
        else:
            # ensemble mode (uq.py:336-339): user-supplied Saltelli output sample
            nb_given = nb                                  # Saltelli sample as written by the driver (uq.py:336-339)
            size = len(self.space) // nb_given
            Y = torch.from_numpy(np.ascontiguousarray(self.output[:nb_given * size], dtype=np.float64)).to(dev)
            if block:
                Y = ((Y[:, 1:] + Y[:, :-1]) / 2.0).sum(dim=1, keepdim=True).contiguous()
            shift_f = Y[:size].mean(dim=0).contiguous()
            _lib.check(lib.bk_sobol_accumulate_y(ptr(Y), size, size, p, F, second, ptr(shift_f), ptr(partials),
                                                 stream))

        totals = torch.empty((F, nsums, 2), dtype=torch.float64, device=dev)
        _lib.check(lib.bk_sobol_reduce_slots(ptr(partials), F, nsums, ptr(totals), stream))
        if world > 1:
            allreduce_totals(totals)
        out = torch.empty(lib.bk_sobol_result_len(p, F), dtype=torch.float64, device=dev)
        _lib.check(lib.bk_sobol_finalize(ptr(totals), size, p, F, second, ptr(out), stream))
        res = out.cpu().numpy()
        self.details = self._unpack(res, p, F)
        d = self.details
        self._intervals(lib, partials, totals, size, p, F, second, nslots, world, stream)
        self.logger.debug("-> Second order:\n{}\n".format(d['S2']))
        self.logger.debug("-> First order:\n{}\n-> Total:\n{}\n".format(d['S1'], d['ST']))
        if self.type_indices == 'aggregated':
            aggregated = [d['S2_agg'], d['S1_agg'], d['ST_agg']]
            self.logger.info("-> First order confidence:\n{}\n-> Total order confidence:\n{}\n"
                             .format(d['S1_interval'], d['ST_interval']))
            self.logger.info("Aggregated_indices:\n-> Second order:\n{}\n-> First order:\n{}\n-> Total order:\n{}\n"
                             .format(*aggregated))
        else:
            aggregated = [d['S2'][0], d['S1'][0], d['ST'][0]]            # uq.py:466-468
        self._write_files(d, p, F)
        if self.type_indices != 'aggregated':
            self.xdata = None
        return aggregated

    def _intervals(self, lib, partials, totals, size, p, F, second, nslots, world, stream):
        """95 % intervals of the aggregated first / total order indices (uq.py:341, 425-426).

        The reference asks OpenTURNS for asymptotic (delta-method) intervals; those formulas are not pinned in
        this environment, so the drop-in uses a delete-a-group jackknife over the partial-sum slots (disjoint
        groups of base rows, every rank's slots together): asymptotically equivalent, no second pass over the
        design.  ``bk_sobol_jackknife`` evaluates the estimator without each group on the device; the few hundred
        leave-one-out values are combined here (group-size weighted, oracle/saltelli.py:jackknife_aggregated)."""
        import torch
        import torch.distributed as dist
        d = self.details
        jk = torch.empty((nslots, 2 + 2 * p), dtype=torch.float64, device=partials.device)
        _lib.check(lib.bk_sobol_jackknife(ctypes.c_void_p(partials.data_ptr()), ctypes.c_void_p(totals.data_ptr()),
                                          size, p, F, second, ctypes.c_void_p(jk.data_ptr()), stream))
        if world > 1:
            parts = [torch.empty_like(jk) for _ in range(world)]
            dist.all_gather(parts, jk)
            jk = torch.cat(parts, dim=0)
        jk = jk.cpu().numpy()
        jk = jk[jk[:, 0] > 0]
        theta = np.concatenate([d['S1_agg'], d['ST_agg']])
        G = len(jk)
        if G < 2:
            sigma = np.full(2 * p, np.nan)
        else:
            h = (size / jk[:, 1])[:, None]
            th_g = jk[:, 2:]
            theta_j = G * theta - ((1.0 - 1.0 / h) * th_g).sum(axis=0)
            pseudo = h * theta - (h - 1.0) * th_g
            sigma = np.sqrt((((pseudo - theta_j) ** 2) / (h - 1.0)).sum(axis=0) / G)
        z = 1.959963984540054                                            # 95 %, two-sided
        d['S1_agg_sigma'], d['ST_agg_sigma'] = sigma[:p], sigma[p:]
        d['S1_interval'] = (d['S1_agg'] - z * sigma[:p], d['S1_agg'] + z * sigma[:p])
        d['ST_interval'] = (d['ST_agg'] - z * sigma[p:], d['ST_agg'] + z * sigma[p:])
        d['jackknife_groups'] = G

    def _write_files(self, d, p, F):
        """sensitivity.json / sensitivity_aggregated.json as the reference writes them (uq.py:380-393, 437-463;
        layout of input_output/formater.py:61-77: {name: [[value], ...]})."""
        if self.fname is None:
            self.logger.debug("No output folder to write indices in")
            return
        import json
        import os

        def write(path, data, names):
            data = np.atleast_2d(data)
            with open(path, 'w') as fd:
                json.dump({name: data[:, j:j + 1].tolist() for j, name in enumerate(names)}, fd)
        os.makedirs(self.fname, exist_ok=True)
        block = self.type_indices == 'block'
        data = np.append(np.reshape(d['S1'], (F, p)), np.reshape(d['ST'], (F, p)), axis=1)
        names = ['S_{}'.format(lab) for lab in self.plabels] + ['S_T_{}'.format(lab) for lab in self.plabels]
        if (self.output_len > 1) and not block:
            names = ['x'] + names
            data = np.append(np.reshape(self.xdata, (F, 1)), data, axis=1)
        write(os.path.join(self.fname, 'sensitivity.json'), data, names)
        if self.type_indices == 'aggregated':
            (i1_min, i1_max), (i2_min, i2_max) = d['S1_interval'], d['ST_interval']
            agg = np.array([i1_min, d['S1_agg'], i1_max, i2_min, d['ST_agg'], i2_max]).flatten()
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
        return d
