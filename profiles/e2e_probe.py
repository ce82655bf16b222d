"""Host boundary of DeviceModel.predict on this box: raw host copy rates (pageable -> pinned memcpy, pageable -> device
through the driver, pinned -> device) and the end-to-end rate of the staged pipeline against the plain
driver-staged pipeline (pageable tensors handed straight to cudaMemcpyAsync).  One JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from batman_b200 import SurrogateModel, UQ, _lib  # noqa: E402


def driver_staged(model, points, chunk):
    """Round-1 pipeline: pageable slices go to the device with .to(non_blocking=True) (the driver stages them)."""
    k = points.shape[0]
    src = torch.from_numpy(points)
    out_mean, out_std = np.empty((k, model.m)), np.empty((k, model.m))
    s_main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    bounds = [(lo, min(lo + chunk, k)) for lo in range(0, k, chunk)]
    pending = None
    for c in range(len(bounds) + 1):
        nxt = None
        if c < len(bounds):
            lo, hi = bounds[c]
            with torch.cuda.stream(s_in):
                pts = src[lo:hi].to(model.device, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            s_main.wait_event(ev_in)
            pts.record_stream(s_main)
            mean, std = model.predict_device(pts, True, check=False)
            ev = torch.cuda.Event()
            ev.record(s_main)
            nxt = (lo, hi, mean, std, ev)
        if pending is not None:
            plo, phi, pmean, pstd, pev = pending
            with torch.cuda.stream(s_out):
                s_out.wait_event(pev)
                pmean.record_stream(s_out)
                pstd.record_stream(s_out)
                torch.from_numpy(out_mean[plo:phi]).copy_(pmean)
                torch.from_numpy(out_std[plo:phi]).copy_(pstd)
        pending = nxt
    s_main.synchronize()
    return out_mean, out_std


def best(fn, reps=3):
    t = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t.append(time.perf_counter() - t0)
    return min(t)


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
    out = {"rows": rows, "threads": torch.get_num_threads(), "cpus": os.cpu_count()}
    # raw copies of one slice
    n = model.E2E_CHUNK
    pin = torch.empty((n, bench.P_DIM), dtype=torch.float64).pin_memory()
    src = torch.from_numpy(host)
    gb = n * bench.P_DIM * 8 / 1e9
    out["memcpy_pageable_to_pinned_GBs"] = gb / best(lambda: pin.copy_(src[:n]), 5)
    out["memcpy_numpy_GBs"] = gb / best(lambda: np.copyto(pin.numpy(), host[:n]), 5)
    out["h2d_pageable_GBs"] = gb / best(lambda: src[:n].to(dev), 5)
    out["h2d_pinned_GBs"] = gb / best(lambda: pin.to(dev, non_blocking=True), 5)
    fresh = lambda: np.empty((n, 2))
    dres = torch.rand((n, 2), dtype=torch.float64, device=dev)
    out["d2h_fresh_pageable_GBs"] = n * 16 / 1e9 / best(lambda: torch.from_numpy(fresh()).copy_(dres), 5)
    model.predict(host[:4096])
    for name, fn in (("staged", lambda: model.predict(host)), ("driver_staged_2chunks", lambda: driver_staged(model, host, n // 2)),
                     ("driver_staged_4chunks", lambda: driver_staged(model, host, n)),
                     ("staged_again", lambda: model.predict(host))):
        out[name + "_pred_per_s"] = rows / best(fn, 3)
    out["resident_pred_per_s"] = rows / best(lambda: model.predict_device(D, True, check=False), 3)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
