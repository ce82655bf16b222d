"""Host boundary of DeviceModel.predict on this box: raw host copy rates (pageable -> pinned memcpy, pageable -> device
through the driver, pinned -> device) and the end-to-end time of a pinned staging ring ("staged": helper threads copy
slices through pinned buffers; measured, not shipped) against the driver-staged pipeline (pageable tensors handed
straight to cudaMemcpyAsync: what DeviceModel.predict does, "shipped_predict").  One JSON line per process; under
torchrun every rank is an independent replica on its own GPU."""
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


def _staging(self, return_std):
    """Two pinned input and two pinned output slices (allocated once per model): the caller's pageable arrays are
    copied through them by helper threads, so every host<->device transfer is asynchronous."""
    st = getattr(self, '_stage', None)
    if st is None:
        st = {'in': [torch.empty((self.E2E_CHUNK_RING, self.p), dtype=torch.float64).pin_memory() for _ in range(2)],
              'mean': [torch.empty((self.E2E_CHUNK_RING, self.m), dtype=torch.float64).pin_memory() for _ in range(2)],
              'std': None}
        self._stage = st
    if return_std and st['std'] is None:
        st['std'] = [torch.empty((self.E2E_CHUNK_RING, self.m), dtype=torch.float64).pin_memory() for _ in range(2)]
    return st

def pinned_ring(self, points, return_std=True):
    """The pinned staging ring that was measured and NOT shipped: helper threads copy slices of the pageable
    array into two pinned buffers and results out of two pinned buffers."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    k = points.shape[0]
    src = torch.from_numpy(points)
    with torch.cuda.device(self.device):
        if k <= self.E2E_CHUNK_RING:
            pts = src.to(self.device, non_blocking=False)
            mean, std = self.predict_device(pts, return_std)
            return mean.cpu().numpy(), (std.cpu().numpy() if std is not None else None)
        from concurrent.futures import ThreadPoolExecutor
        out_mean = np.empty((k, self.m))
        out_std = np.empty((k, self.m)) if return_std else None
        dst_mean = torch.from_numpy(out_mean)
        dst_std = torch.from_numpy(out_std) if return_std else None
        stage = _staging(self, return_std)
        s_main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        bounds = [(lo, min(lo + self.E2E_CHUNK_RING, k)) for lo in range(0, k, self.E2E_CHUNK_RING)]
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
    local = int(os.environ.get('LOCAL_RANK', '0'))      # under torchrun: independent replicas, one per GPU
    dev = torch.device('cuda', local)
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
    out = {"rank": local, "rows": rows, "threads": torch.get_num_threads(), "cpus": os.cpu_count(),
           "affinity": len(os.sched_getaffinity(0))}
    # raw copies of one slice
    n = 4 * 148 * 128 * 8
    model.E2E_CHUNK_RING = n
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
    for name, fn in (("staged", lambda: pinned_ring(model, host)), ("driver_staged_2chunks", lambda: driver_staged(model, host, n // 2)),
                     ("driver_staged_4chunks", lambda: driver_staged(model, host, n)),
                     ("staged_again", lambda: pinned_ring(model, host)), ("shipped_predict", lambda: model.predict(host))):
        ts = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(round(time.perf_counter() - t0, 4))
        out[name + "_s"] = ts
    out["resident_pred_per_s"] = rows / best(lambda: model.predict_device(D, True, check=False), 3)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
