"""One batched LML call (K5) at a given size, for `ncu --metrics gpu__time_duration.sum` launch lists and wall time.

    python profiles/lml_profile.py n p R [reps]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from batman_b200 import kernel_spec  # noqa: E402
from batman_b200.lml import lml_batched  # noqa: E402


def main():
    n, p, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    X = bench.halton(n, p)
    y = bench.g_function(X)[:, 0]
    kern = ConstantKernel(1.0) * Matern(length_scale=[0.7] * p, nu=1.5)
    rng = np.random.RandomState(1)
    specs = [kernel_spec.flatten(kern.clone_with_theta(kern.theta + rng.uniform(-0.5, 0.5, p + 1)), p) for _ in range(R)]
    dev = torch.device('cuda', 0)
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
    yd = torch.from_numpy((y - y.mean()) / y.std()).to(dev)
    lml_batched(Xd, yd, specs)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals = lml_batched(Xd, yd, specs)
        best = min(best, time.perf_counter() - t0)
    flop = R * (n ** 3 / 3.0 + 2.0 * n ** 2)
    print('{"n": %d, "p": %d, "R": %d, "ms_per_batch": %.3f, "ms_per_lml": %.4f, "algorithmic_tflops": %.2f, "finite": %d}'
          % (n, p, R, best * 1e3, best * 1e3 / R, flop / best * 1e-12, int(np.isfinite(vals).sum())))


if __name__ == '__main__':
    main()
