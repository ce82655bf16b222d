"""K1 || sigma-kernel overlap (bk_gp_set_overlap) on / off: device-resident step time, end-to-end time, clocks and
board power of each mode, same model and design as bench.py (cfg2).  Run under gpurun; one JSON line per mode."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from batman_b200 import SurrogateModel, UQ, _lib  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    cx = bench.Ctx()
    cx.torch, cx.world, cx.rank, cx.dev, cx.lib, cx._lib = torch, 1, 0, dev, lib, _lib
    cx.SurrogateModel, cx.UQ = SurrogateModel, UQ
    X, y = bench.training_set()
    corners = np.array([[0.0] * bench.P_DIM, [1.0] * bench.P_DIM])
    surrogate = bench._fixed_theta_surrogate(cx, X, y, corners, bench.fixed_kernel())
    model = surrogate.predictor._model()
    model.ensure_w()
    rows = 18_000_000
    D = torch.rand((rows, bench.P_DIM), dtype=torch.float64, device=dev)
    host = D.cpu().numpy()
    for on in (0, 1, 0, 1):
        _lib.check(lib.bk_gp_set_overlap(model.handle, on))
        for _ in range(2):
            model.predict_device(D, return_std=True, check=False)
        torch.cuda.synchronize()
        sampler = bench.ClockSampler(0)
        sampler.start()
        lib.bk_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps = 8
        for _ in range(steps):
            model.predict_device(D, return_std=True, check=False)
        e1.record()
        torch.cuda.synchronize()
        lib.bk_profile_enable(0)
        ms_arr, cnt_arr = (ctypes.c_double * 8)(), (ctypes.c_long * 8)()
        _lib.check(lib.bk_profile_read(ms_arr, cnt_arr))
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1) / steps
        surrogate(host[:4096])
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            surrogate(host)
            t.append(time.perf_counter() - t0)
        print(json.dumps({"overlap": on, "ms_per_step": ms, "pred_per_s": rows / ms * 1e3,
                          "k1_avg_ms": ms_arr[0] / max(1, cnt_arr[0]), "sigma_avg_ms": ms_arr[1] / max(1, cnt_arr[1]),
                          "e2e_pred_per_s": rows / min(t), "clocks": clocks}), flush=True)


if __name__ == '__main__':
    main()
