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

    E2E_CHUNK = 4 * 148 * 128 * 8     # rows per pipelined host slice (four K* chunks of 151 552 rows)

    def _staging(self, return_std):
        """Two pinned input and two pinned output slices (allocated once per model): the caller's pageable arrays are
        copied through them by helper threads, so every host<->device transfer is asynchronous."""
        st = getattr(self, '_stage', None)
        if st is None:
            st = {'in': [torch.empty((self.E2E_CHUNK, self.p), dtype=torch.float64).pin_memory() for _ in range(2)],
                  'mean': [torch.empty((self.E2E_CHUNK, self.m), dtype=torch.float64).pin_memory() for _ in range(2)],
                  'std': None}
            self._stage = st
        if return_std and st['std'] is None:
            st['std'] = [torch.empty((self.E2E_CHUNK, self.m), dtype=torch.float64).pin_memory() for _ in range(2)]
        return st

    def predict(self, points, return_std=True):
        """numpy in, numpy out (the Kriging.evaluate boundary): host<->device copies included.

        Large batches are pipelined over slices of E2E_CHUNK rows.  A helper thread copies slice c+1 of the caller's
        (pageable) array into a pinned staging buffer while the GPU predicts slice c; the upload, the kernels and the
        download of consecutive slices run on three streams; a second helper thread moves finished results from the
        pinned output buffer into the returned arrays.  The GPU never waits for a pageable copy."""
        points = np.ascontiguousarray(points, dtype=np.float64)
        k = points.shape[0]
        src = torch.from_numpy(points)
        with torch.cuda.device(self.device):
            if k <= self.E2E_CHUNK:
                pts = src.to(self.device, non_blocking=False)
                mean, std = self.predict_device(pts, return_std)
                return mean.cpu().numpy(), (std.cpu().numpy() if std is not None else None)
            from concurrent.futures import ThreadPoolExecutor
            out_mean = np.empty((k, self.m))
            out_std = np.empty((k, self.m)) if return_std else None
            dst_mean = torch.from_numpy(out_mean)
            dst_std = torch.from_numpy(out_std) if return_std else None
            stage = self._staging(return_std)
            s_main = torch.cuda.current_stream()
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            bounds = [(lo, min(lo + self.E2E_CHUNK, k)) for lo in range(0, k, self.E2E_CHUNK)]
            n_sl = len(bounds)
            up_done = [None] * n_sl          # events: upload of slice c has left its pinned buffer
            down_done = [None] * n_sl        # events: results of slice c are in their pinned buffer
            out_futs = [None] * n_sl

            def stage_in(c):
                lo, hi = bounds[c]
                if c >= 2:
                    up_done[c - 2].synchronize()         # buffer c & 1 was last uploaded for slice c - 2
                stage['in'][c & 1][:hi - lo].copy_(src[lo:hi])

            def stage_out(c):
                lo, hi = bounds[c]
                down_done[c].synchronize()
                dst_mean[lo:hi].copy_(stage['mean'][c & 1][:hi - lo])
                if return_std:
                    dst_std[lo:hi].copy_(stage['std'][c & 1][:hi - lo])

            with ThreadPoolExecutor(1) as pool_in, ThreadPoolExecutor(1) as pool_out:
                fut_in = pool_in.submit(stage_in, 0)
                for c, (lo, hi) in enumerate(bounds):
                    fut_in.result()
                    with torch.cuda.stream(s_in):
                        pts = stage['in'][c & 1][:hi - lo].to(self.device, non_blocking=True)
                        up_done[c] = torch.cuda.Event()
                        up_done[c].record(s_in)
                    if c + 1 < n_sl:
                        fut_in = pool_in.submit(stage_in, c + 1)
                    s_main.wait_event(up_done[c])
                    pts.record_stream(s_main)
                    mean, std = self.predict_device(pts, return_std, check=False)
                    ev = torch.cuda.Event()
                    ev.record(s_main)
                    if c >= 2:
                        out_futs[c - 2].result()         # output buffer c & 1 has been emptied
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev)
                        mean.record_stream(s_out)
                        stage['mean'][c & 1][:hi - lo].copy_(mean, non_blocking=True)
                        if std is not None:
                            std.record_stream(s_out)
                            stage['std'][c & 1][:hi - lo].copy_(std, non_blocking=True)
                        down_done[c] = torch.cuda.Event()
                        down_done[c].record(s_out)
                    out_futs[c] = pool_out.submit(stage_out, c)
                for f in out_futs:
                    f.result()
            s_main.synchronize()
            if return_std:
                self.check_pipeline_flag()
        return out_mean, out_std
