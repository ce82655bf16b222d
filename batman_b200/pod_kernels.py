"""POD mode -> field step (pod/pod.py:226-228): pred = mean_snapshot + (U @ modes.T).T

Hand-written DMMA kernel (csrc/pod.cu, ``bk_pod_inverse``); the variant fused with the
per-feature Saltelli sums is ``bk_pod_sobol_accumulate`` (used by ``UQ.sobol``).
"""
import ctypes

import torch

from . import _lib


def pod_inverse(mean_dev, U_dev, modes_dev):
    """mean (F,), U (F, m), modes (k, m) CUDA float64 tensors -> (k, F)."""
    lib = _lib.load()
    modes_dev = modes_dev.contiguous()
    k, m = modes_dev.shape
    F = U_dev.shape[0]
    out = torch.empty((k, F), dtype=torch.float64, device=modes_dev.device)
    with torch.cuda.device(modes_dev.device):
        _lib.check(lib.bk_pod_inverse(ctypes.c_void_p(modes_dev.data_ptr()), k, m, ctypes.c_void_p(U_dev.data_ptr()),
                                      ctypes.c_void_p(mean_dev.data_ptr()), F, ctypes.c_void_p(out.data_ptr()),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out
