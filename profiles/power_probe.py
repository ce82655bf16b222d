"""Power / clock of the hot kernels run alone in 4 s loops (nvidia-smi sampled every 100 ms).

    python profiles/power_probe.py > gpurun_out/power_probe.jsonl

Why: bench.py's timed region sits at the 1 kW power cap (sw_power_cap, ~1660 of 1965 MHz).  Whether overlapping the
FP64-pipe kernel (k_predict) with the tensor-pipe kernel (k_var_ozaki) can shorten the step depends on how far below
the cap each of them runs on its own: under a power cap the step time is bounded below by energy / 1 kW.
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from batman_b200 import SurrogateModel, UQ  # noqa: E402

cx = bench.Ctx()
cx.SurrogateModel = SurrogateModel
X, y = bench.training_set()
s = bench._fixed_theta_surrogate(cx, X, y, np.array([[0.0] * 8, [1.0] * 8]), bench.fixed_kernel())
model = s.predictor._model()
model.ensure_w()
pts = torch.rand((148 * 128 * 8 * 4, 8), dtype=torch.float64, device='cuda')
uq = UQ(s, dists=['Uniform(0., 1.)'] * 8, nsample=1000000, intervals='jackknife')


def sample(fn, seconds=4.0):
    fd, path = tempfile.mkstemp(suffix='.csv')
    os.close(fd)
    fn()
    torch.cuda.synchronize()
    proc = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap',
                             '--format=csv,noheader,nounits', '-lms', '100'], stdout=open(path, 'w'))
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        fn()
        n += 1
        if n % 4 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    proc.terminate()
    proc.wait()
    rows = [line.split(',') for line in open(path)]
    os.unlink(path)
    rows = rows[5:-2] if len(rows) > 12 else rows
    mhz = [float(r[0]) for r in rows]
    pw = [float(r[1]) for r in rows]
    capped = sum(1 for r in rows if 'Active' in r[2] and 'Not' not in r[2])
    return dict(calls=n, ms_per_call=1e3 * wall / n, sm_mhz_median=float(np.median(mhz)), power_w_median=float(np.median(pw)),
                power_w_max=float(max(pw)), power_cap_samples=capped, samples=len(rows))


cases = [
    ("k_predict alone (mean only, FP64 pipe)", lambda: model.predict_device(pts, False)),
    ("k_predict + k_var_ozaki (mean + sigma, int8 tensor path)", lambda: model.predict_device(pts, True, check=False)),
    ("k_sobol_static (fused Saltelli, FP64 pipe)", uq.sobol),
]
for name, fn in cases:
    print(json.dumps(dict(case=name, **sample(fn))), flush=True)
model.set_sigma_kernel('dmma')
print(json.dumps(dict(case="k_predict + k_var_trmm (mean + sigma, FP64 DMMA path)",
                      **sample(lambda: model.predict_device(pts, True, check=False)))), flush=True)
time.sleep(3.0)
print(json.dumps(dict(case="idle", **sample(lambda: time.sleep(0.05), seconds=2.0))), flush=True)
