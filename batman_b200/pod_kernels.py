"""POD mode -> field step (pod/pod.py:226-228): pred = mean_snapshot + (U @ modes.T).T

(k, m) x (m, F) with m <= ~20 modes and F ~ 400 features: a genuine GEMM.
Round-1 state: dispatched to the library FP64 GEMM (torch.addmm -> cuBLAS
DGEMM, which runs on the same DMMA pipe); the hand-written DMMA kernel that
fuses this step with the per-feature Saltelli sums is K4 in DESIGN.md.
"""
import torch


def pod_inverse(mean_dev, U_dev, modes_dev):
    return torch.addmm(mean_dev[None, :].expand(modes_dev.shape[0], -1), modes_dev, U_dev.t())
