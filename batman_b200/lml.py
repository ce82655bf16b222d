"""Batched log-marginal-likelihood and final factorisation on the device (K5, csrc/chol.cu).

For R candidate hyper-parameter vectors at once (sklearn:_gpr.py:584-617):
    K_r = k_theta_r(X, X) + 1e-10 I ;  L_r = chol(K_r) ;  z_r = L_r^-1 y
    LML_r = -1/2 ||z_r||^2 - sum(log diag L_r) - n/2 log(2 pi)     (-inf if the Cholesky fails)
All arithmetic runs in ``bk_lml_batched`` / ``bk_gp_factorise`` (hand-written blocked Cholesky on the
FP64 tensor pipe); torch only owns the buffers.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_MAX_BATCH_BYTES = 24 << 30      # matrices of one batch; candidates beyond that are processed in slices


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_ws_cache = {}


def _workspace(device, nbytes):
    buf = _ws_cache.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[device] = buf
    return buf


def lml_batched(X_dev, y_dev, specs, targets=None):
    """X_dev (n, p); y_dev (n,) or (T, n) CUDA float64 tensors (normalised targets of T outputs); specs: list of
    KernelSpec; targets[r] = row of y_dev candidate r is scored on (default 0) -> numpy (R,) LML values.

    One device call per slice of candidates that fits the matrix budget: the candidates of every restart and every
    output of a generation go in together (fit.py)."""
    lib = _lib.load()
    X_dev = X_dev.contiguous()
    Y = (y_dev if y_dev.dim() == 2 else y_dev[None, :]).contiguous()   # the C ABI takes dense row-major buffers
    n, p = X_dev.shape
    R = len(specs)
    targets = np.zeros(R, dtype=np.int64) if targets is None else np.asarray(targets, dtype=np.int64)
    if len(targets) != R or (R and (targets.min() < 0 or targets.max() >= Y.shape[0])):
        raise ValueError("lml_batched: targets must hold one row index of y_dev per candidate")
    out = np.empty(R)
    n_pad = lib.bk_gp_padded_n(n)
    per = max(1, int(_MAX_BATCH_BYTES // (n_pad * n_pad * 8)))
    with torch.cuda.device(X_dev.device):
        for lo in range(0, R, per):
            sl = specs[lo:lo + per]
            r = len(sl)
            uniq, local = np.unique(targets[lo:lo + r], return_inverse=True)     # at most r distinct rows
            y_sl = Y if len(uniq) == Y.shape[0] else Y[torch.from_numpy(uniq).to(Y.device)].contiguous()
            tgt = _lib.as_int_array(local)
            ws = _workspace(X_dev.device, lib.bk_lml_workspace(n, r, 0))
            lml = torch.empty(r, dtype=torch.float64, device=X_dev.device)
            if sl[0].general is not None:
                # candidates of one fit share the tree shape (clone_with_theta only moves the values)
                g0 = sl[0].general
                if any(s.general is None or s.general.shape_key() != g0.shape_key() for s in sl):
                    raise ValueError("lml_batched: candidates of one batch must share the kernel tree")
                _lib.check(lib.bk_lml_batched_general(
                    _ptr(X_dev), _ptr(y_sl), len(uniq), tgt, n, p, r, g0.nf, _lib.as_int_array(g0.fkind), g0.nt,
                    _lib.as_int_array(g0.texp.ravel()),
                    _lib.as_double_array(np.concatenate([s.general.tcoef for s in sl])),
                    _lib.as_double_array(np.concatenate([s.general.fls.ravel() for s in sl])),
                    _lib.as_double_array([s.kdiag for s in sl]), _ptr(lml), _ptr(ws), ws.numel(), _stream()))
                out[lo:lo + r] = lml.cpu().numpy()
                continue
            if any(s.general is not None for s in sl):
                raise ValueError("lml_batched: candidates of one batch must share the kernel tree")
            kinds = _lib.as_int_array([s.kind for s in sl])
            c0 = _lib.as_double_array([s.c0 for s in sl])
            c1 = _lib.as_double_array([s.c1 for s in sl])
            kd = _lib.as_double_array([s.kdiag for s in sl])
            ls = _lib.as_double_array(np.concatenate([np.asarray(s.ls, dtype=np.float64) for s in sl]))
            _lib.check(lib.bk_lml_batched(_ptr(X_dev), _ptr(y_sl), len(uniq), tgt, n, p, r, kinds, c0, c1, kd, ls,
                                          _ptr(lml), _ptr(ws), ws.numel(), _stream()))
            out[lo:lo + r] = lml.cpu().numpy()
    return out


def int8_scheme(n, specs):
    """Slicing scheme of the int8 sigma kernel for a model: 1 (7 x 8-bit, fewer MMAs) when the int32 accumulation
    stays exact (n_pad <= 8192) and every covariance is non-negative (K* is sliced into unsigned bytes), else 0."""
    n_pad = _lib.load().bk_gp_padded_n(n)
    return 1 if n_pad <= 8192 and all(s.nonneg for s in specs) else 0


def factorise(X_dev, y_dev, spec):
    """Final state at the chosen theta (sklearn:_gpr.py:349-367).

    -> (L (n, n) numpy lower, alpha (n,) numpy, lml float, (FP64 tiles, int8 slices, row scales, slicing scheme)
    of L^-1 as CUDA tensors)."""
    lib = _lib.load()
    X_dev, y_dev = X_dev.contiguous(), y_dev.contiguous()
    n, p = X_dev.shape
    n_pad = lib.bk_gp_padded_n(n)
    dev = X_dev.device
    scheme = int8_scheme(n, [spec])
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.bk_lml_workspace(n, 1, 1))
        L = torch.empty((n_pad, n_pad), dtype=torch.float64, device=dev)
        alpha = torch.empty(n_pad, dtype=torch.float64, device=dev)
        wt = torch.empty(lib.bk_gp_wtiles_len(n), dtype=torch.float64, device=dev)
        wq = torch.empty(lib.bk_gp_wq_bytes(n), dtype=torch.int8, device=dev)
        wsig = torch.empty(n_pad, dtype=torch.float64, device=dev)
        lml = torch.empty(1, dtype=torch.float64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        if spec.general is not None:
            g = spec.general
            _lib.check(lib.bk_gp_factorise_general(
                _ptr(X_dev), _ptr(y_dev), n, p, g.nf, _lib.as_int_array(g.fkind), g.nt,
                _lib.as_int_array(g.texp.ravel()), _lib.as_double_array(g.tcoef), _lib.as_double_array(g.fls.ravel()),
                spec.kdiag, _ptr(L), _ptr(alpha), _ptr(wt), _ptr(wq), _ptr(wsig), scheme, _ptr(lml),
                _ptr(info), _ptr(ws), ws.numel(), _stream()))
        else:
            _lib.check(lib.bk_gp_factorise(_ptr(X_dev), _ptr(y_dev), n, p, spec.kind, spec.c0, spec.c1, spec.kdiag,
                                           _lib.as_double_array(spec.ls), _ptr(L), _ptr(alpha), _ptr(wt), _ptr(wq),
                                           _ptr(wsig), scheme, _ptr(lml), _ptr(info), _ptr(ws),
                                           ws.numel(), _stream()))
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError(
                "The kernel is not returning a positive definite matrix. Try gradually increasing the "
                "'alpha' parameter of your GaussianProcessRegressor estimator.")   # sklearn:_gpr.py:354-360
        L_host = np.tril(L[:n, :n].cpu().numpy())
        return L_host, alpha[:n].cpu().numpy(), float(lml.item()), (wt, wq, wsig, scheme)


def inverse_tiles(L_host, device, scheme=0):
    """L^-1 from a given Cholesky factor (models adopted from scikit-learn state): FP64 DMMA tiles, int8 slices
    and their row scales."""
    lib = _lib.load()
    n = L_host.shape[0]
    with torch.cuda.device(device):
        L = torch.from_numpy(np.ascontiguousarray(L_host, dtype=np.float64)).to(device)
        ws = _workspace(device, lib.bk_lml_workspace(n, 1, 1))
        wt = torch.empty(lib.bk_gp_wtiles_len(n), dtype=torch.float64, device=device)
        wq = torch.empty(lib.bk_gp_wq_bytes(n), dtype=torch.int8, device=device)
        wsig = torch.empty(lib.bk_gp_padded_n(n), dtype=torch.float64, device=device)
        _lib.check(lib.bk_tri_inverse_pack(_ptr(L), n, _ptr(wt), _ptr(wq), _ptr(wsig), scheme, _ptr(ws), ws.numel(),
                                           _stream()))
        torch.cuda.current_stream().synchronize()
    return wt, wq, wsig
