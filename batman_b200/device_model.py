"""Device-resident state of a fitted Kriging model (m independent GPs, one design).

Holds torch CUDA tensors (plumbing: device memory + streams) and the opaque
C handle; every computation goes through the C ABI.  Data layout in HBM, per
output q (DESIGN.md §"Data layout"):

* ``xs[q]``     (n_pad, P) float64   X_train / length_scale, zero padded
                (general kernel trees: (n_pad, nf*P), one scaled copy per stationary leaf)
* ``alpha[q]``  (n_pad,)   float64   K^-1 y (sklearn ``alpha_``), zero padded
* ``wt[q]``     (n_pad*n_pad,) float64  L^-1 packed in DMMA-fragment tiles
                (built lazily: only sigma needs it)
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def current_stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceModel:
    def __init__(self, X, specs, alphas, Ls, y_means, y_stds, device=None, wts=None):
        if not torch.cuda.is_available():
            raise RuntimeError("batman_b200 needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.n, self.p = X.shape
        self.m = len(specs)
        self.P = self.lib.bk_gp_padded_dim(self.p)
        if self.P < 0:
            raise NotImplementedError("input dimension %d has no instantiated device kernel (max 32)" % self.p)
        self.n_pad = self.lib.bk_gp_padded_n(self.n)
        self.specs = specs
        self._Ls = Ls                    # host copies; W is built from them on first sigma request
        self.y_means = [float(v) for v in y_means]
        self.y_stds = [float(v) for v in y_stds]
        handle = ctypes.c_void_p()
        _lib.check(self.lib.bk_gp_create(ctypes.byref(handle), self.n, self.p, self.m))
        self.handle = handle
        with torch.cuda.device(self.device):
            Xd = torch.from_numpy(X).to(self.device)
            self.xs, self.alpha = [], []
            # wts[q] = (FP64 tiles, int8 slices, row scales, slicing scheme) of L^-1 when the fit already produced
            # them.  The sigma kernel uses ONE slicing per model: an output whose fit sliced with the other scheme
            # (its own coefficients allowed unsigned bytes, another output's do not) is re-sliced by ensure_w().
            from .lml import int8_scheme
            self.int8_scheme = int8_scheme(self.n, specs)
            self.wt, self.wq, self.wsig = [None] * self.m, [None] * self.m, [None] * self.m
            for q, w in enumerate(wts or []):
                if w is not None and w[3] == self.int8_scheme:
                    self.wt[q], self.wq[q], self.wsig[q] = w[0], w[1], w[2]
            for q, spec in enumerate(specs):
                if spec.general is not None:
                    g = spec.general
                    if g.nf * self.P > 48:
                        raise NotImplementedError(
                            "kernel tree with %d stationary leaves at input dimension %d exceeds the staging "
                            "buffers of the general device kernel (leaves x padded dimension <= 48)" % (g.nf, self.p))
                    xs = torch.zeros((self.n_pad, g.nf * self.P), dtype=torch.float64, device=self.device)
                    for f in range(g.nf):
                        xs[:self.n, f * self.P:f * self.P + self.p] = Xd / torch.from_numpy(g.fls[f]).to(self.device)
                else:
                    ls = torch.from_numpy(np.asarray(spec.ls, dtype=np.float64)).to(self.device)
                    xs = torch.zeros((self.n_pad, self.P), dtype=torch.float64, device=self.device)
                    xs[:self.n, :self.p] = Xd / ls        # IEEE division, same bits as numpy's X / length_scale
                al = torch.zeros(self.n_pad, dtype=torch.float64, device=self.device)
                al[:self.n] = torch.from_numpy(np.ascontiguousarray(alphas[q], dtype=np.float64).ravel()).to(self.device)
                self.xs.append(xs)
                self.alpha.append(al)
                self._set_output(q)
        self._workspace = None
        self._scaler = None
        self.sigma_kernel = 'dmma'
        # ONE sigma path per model size, no environment switches: the int8 tensor-core kernel (exact slicing,
        # float64-equivalent results) wherever its int32 accumulation is exact (n_train <= 16384), the FP64 DMMA
        # kernel above.  set_sigma_kernel / set_int8_scheme remain as explicit calls for the parity tests, which
        # check both kernels and both slicings against each other.
        self.set_sigma_kernel('int8' if self.n_pad <= 16384 else 'dmma')

    def set_sigma_kernel(self, name):
        """'dmma' = FP64 tensor cores (k_var_trmm); 'int8' = tcgen05 int8 + TMEM, error-free slicing (k_var_ozaki;
        8 x 7-bit or 7 x 8-bit slices, see lml.int8_scheme)."""
        if name not in ('dmma', 'int8'):
            raise ValueError("sigma kernel must be 'dmma' or 'int8'")
        _lib.check(self.lib.bk_gp_set_var_mode(self.handle, 1 + self.int8_scheme if name == 'int8' else 0))
        self.sigma_kernel = name

    def set_int8_scheme(self, scheme):
        """Force the slicing scheme of the int8 sigma kernel (0: 8 x 7-bit, 1: 7 x 8-bit); L^-1 is re-sliced."""
        if scheme not in (0, 1):
            raise ValueError("int8 scheme must be 0 or 1")
        if self._Ls is None:
            raise RuntimeError("re-slicing needs the Cholesky factor")
        self.int8_scheme = scheme
        self.wt, self.wq, self.wsig = [None] * self.m, [None] * self.m, [None] * self.m
        self.ensure_w()
        self.set_sigma_kernel(self.sigma_kernel)

    # -- C handle ------------------------------------------------------------------------------
    def _set_output(self, q):
        spec = self.specs[q]
        if spec.general is not None:
            g = spec.general
            _lib.check(self.lib.bk_gp_set_output_general(
                self.handle, q, g.nf, _lib.as_int_array(g.fkind), _lib.as_double_array(g.fls.ravel()), g.nt,
                _lib.as_double_array(g.tcoef), _lib.as_int_array(g.texp.ravel()), spec.kdiag,
                _ptr(self.xs[q]), _ptr(self.alpha[q]), _ptr(self.wt[q]), self.y_means[q], self.y_stds[q]))
        else:
            _lib.check(self.lib.bk_gp_set_output(
                self.handle, q, spec.kind, spec.c0, spec.c1, spec.kdiag, _lib.as_double_array(spec.ls),
                _ptr(self.xs[q]), _ptr(self.alpha[q]), _ptr(self.wt[q]), self.y_means[q], self.y_stds[q]))
        _lib.check(self.lib.bk_gp_set_output_int8(self.handle, q, _ptr(self.wq[q]), _ptr(self.wsig[q])))

    def set_scaler(self, scale, mn):
        """MinMaxScaler fused into the kernels (surrogate_model.py:181); None = identity."""
        if scale is None:
            _lib.check(self.lib.bk_gp_set_scaler(self.handle, None, None))
        else:
            _lib.check(self.lib.bk_gp_set_scaler(self.handle, _lib.as_double_array(scale), _lib.as_double_array(mn)))
        self._scaler = (scale, mn)

    def ensure_w(self):
        """W = L^-1 (once per model) as packed DMMA fragment tiles: blocked triangular inversion on the
        device (bk_tri_inverse_pack), unless the fit already produced the tiles."""
        if all(w is not None for w in self.wt):
            return
        if self._Ls is None:
            raise RuntimeError("sigma requested but the model holds no Cholesky factor")
        from .lml import inverse_tiles
        for q in range(self.m):
            if self.wt[q] is None:
                self.wt[q], self.wq[q], self.wsig[q] = inverse_tiles(self._Ls[q], self.device, self.int8_scheme)
                self._set_output(q)

    def close(self):
        if getattr(self, 'handle', None) is not None and self.handle.value:
            self.lib.bk_gp_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- compute -------------------------------------------------------------------------------
    def workspace(self, k, want_std):
        need = self.lib.bk_gp_predict_workspace(self.handle, k, 1 if want_std else 0)
        if need == 0:
            return None, 0
        if self._workspace is None or self._workspace.numel() * 8 < need:
            self._workspace = torch.empty((need + 7) // 8, dtype=torch.float64, device=self.device)
        return self._workspace, self._workspace.numel() * 8

    def check_pipeline_flag(self):
        """k_var_ozaki records an mbarrier time-out (a pipeline fault) instead of hanging the GPU;
        ``bk_gp_predict_status`` reads the flag back (it synchronises the stream, so the pipelined host path
        checks once per call)."""
        ws = self._workspace
        if ws is None or self.sigma_kernel != 'int8':
            return
        fault = ctypes.c_int(0)
        _lib.check(self.lib.bk_gp_predict_status(_ptr(ws), ws.numel() * 8, ctypes.byref(fault), current_stream_ptr()))
        if fault.value != 0:
            raise _lib.BatmanB200Error("k_var_ozaki: an mbarrier wait timed out (pipeline fault); sigma is invalid")

    def predict_device(self, pts_dev, return_std=True, check=True):
        """pts_dev: (k, p) float64 CUDA tensor -> (mean, std) CUDA tensors (k, m); std None if not requested."""
        pts_dev = pts_dev.contiguous()          # the C ABI takes dense row-major buffers
        k = pts_dev.shape[0]
        mean = torch.empty((k, self.m), dtype=torch.float64, device=self.device)
        std = torch.empty((k, self.m), dtype=torch.float64, device=self.device) if return_std else None
        if k == 0:
            return mean, std
        if return_std:
            self.ensure_w()
        ws, ws_bytes = self.workspace(k, return_std)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bk_gp_predict(self.handle, _ptr(pts_dev), k, _ptr(mean), _ptr(std), _ptr(ws),
                                              ws_bytes, current_stream_ptr()))
            if return_std and check:
                self.check_pipeline_flag()
        return mean, std

    E2E_CHUNK = 2 * 148 * 128 * 8     # rows per pipelined host slice (two K* chunks of 151 552 rows)

    def predict(self, points, return_std=True):
        """numpy in, numpy out (the Kriging.evaluate boundary): host<->device copies included.

        Large batches are pipelined: while slice c is being predicted, slice c+1 is uploaded on a copy stream
        and the results of slice c-1 are downloaded on another (the download blocks the host thread only, the
        GPU already has the next slice's kernels queued).  Pageable arrays are handed to the driver as they are: its
        staged copies (10-20 GB/s) were steadier than a pinned staging ring filled by helper threads, alone and with
        4 ranks on one host (profiles/r02_e2e_probe*.jsonl).  Also measured and not kept: downloads on a helper
        thread (0.325 s per 1.8e7 rows against 0.285 s inline) and result arrays on madvise(MADV_HUGEPAGE) mappings
        (0.43 s: huge-page faults compact memory on these hosts)."""
        points = np.ascontiguousarray(points, dtype=np.float64)
        k = points.shape[0]
        src = torch.from_numpy(points)
        with torch.cuda.device(self.device):
            if k <= self.E2E_CHUNK:
                pts = src.to(self.device, non_blocking=False)
                mean, std = self.predict_device(pts, return_std)
                return mean.cpu().numpy(), (std.cpu().numpy() if std is not None else None)
            out_mean = np.empty((k, self.m))
            out_std = np.empty((k, self.m)) if return_std else None
            s_main = torch.cuda.current_stream()
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            bounds = [(lo, min(lo + self.E2E_CHUNK, k)) for lo in range(0, k, self.E2E_CHUNK)]
            pending = None
            for c in range(len(bounds) + 1):
                if c < len(bounds):
                    lo, hi = bounds[c]
                    with torch.cuda.stream(s_in):
                        pts = src[lo:hi].to(self.device, non_blocking=True)
                        ev_in = torch.cuda.Event()
                        ev_in.record(s_in)
                    s_main.wait_event(ev_in)
                    pts.record_stream(s_main)
                    mean, std = self.predict_device(pts, return_std, check=False)
                    ev = torch.cuda.Event()
                    ev.record(s_main)
                    nxt = (lo, hi, mean, std, ev)
                else:
                    nxt = None
                if pending is not None:
                    plo, phi, pmean, pstd, pev = pending
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(pev)
                        pmean.record_stream(s_out)
                        torch.from_numpy(out_mean[plo:phi]).copy_(pmean)
                        if pstd is not None:
                            pstd.record_stream(s_out)
                            torch.from_numpy(out_std[plo:phi]).copy_(pstd)
                pending = nxt
            s_main.synchronize()
            if return_std:
                self.check_pipeline_flag()
        return out_mean, out_std
