"""Fixed cost of one UQ.sobol() call: wall time against the base size N on one GPU (cfg2 model), and where the host
time of a small call goes (cProfile, top entries)."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from batman_b200 import SurrogateModel, UQ, _lib  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    cx = bench.Ctx()
    cx.torch, cx.world, cx.rank, cx.dev, cx.lib, cx._lib = torch, 1, 0, dev, _lib.load(), _lib
    cx.SurrogateModel, cx.UQ = SurrogateModel, UQ
    X, y = bench.training_set()
    corners = np.array([[0.0] * bench.P_DIM, [1.0] * bench.P_DIM])
    s = bench._fixed_theta_surrogate(cx, X, y, corners, bench.fixed_kernel())
    dists = ['Uniform(0., 1.)'] * bench.P_DIM
    out = {}
    for N in (1000, 10000, 100000, 1000000):
        for kw, tag in (({}, 'default'), ({'intervals': 'jackknife'}, 'jackknife_only')):
            uq = UQ(s, dists=dists, nsample=N, seed=bench.SEED, **kw)
            uq.sobol()
            t = []
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                uq.sobol()
                torch.cuda.synchronize()
                t.append(time.perf_counter() - t0)
            out['N=%d %s ms' % (N, tag)] = round(min(t) * 1e3, 3)
    print(json.dumps(out))
    uq = UQ(s, dists=dists, nsample=1000, seed=bench.SEED)
    uq.sobol()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        uq.sobol()
    pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats('cumulative').print_stats(28)
    print(buf.getvalue())


if __name__ == '__main__':
    main()
