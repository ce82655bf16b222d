"""Throughput of Kriging mean+sigma prediction across model sizes (BASELINE configs 2, 3, 5 shapes), 1 GPU.

    python profiles/sweep_configs.py > gpurun_out/sweep.jsonl

Models are built at a fixed theta by the product's own device factorisation (no optimiser); points are
resident in HBM; time = CUDA events over 3 repetitions after 1 warm-up.  Also times the batched LML (K5)
for a DE-generation-sized batch, and checks a few predictions against scikit-learn on the host.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.stats import qmc  # noqa: E402
from sklearn.gaussian_process import GaussianProcessRegressor  # noqa: E402
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern  # noqa: E402

from batman_b200 import Kriging, kernel_spec  # noqa: E402
from batman_b200.fit import normalise_y  # noqa: E402
from batman_b200.lml import lml_batched  # noqa: E402

CASES = [
    dict(name="cfg2 G_Function 8-D n=1024 Matern32", n=1024, p=8, kind='m32', ls=1.5, k=1 << 21, R=135),
    dict(name="cfg3 Michalewicz 10-D n=4096 Matern32", n=4096, p=10, kind='m32', ls=1.0, k=1 << 19, R=32),
    dict(name="north-star 8-D n=4096 Matern32", n=4096, p=8, kind='m32', ls=1.0, k=1 << 19, R=32),
    dict(name="cfg5 synthetic 20-D n=16384 RBF", n=16384, p=20, kind='rbf', ls=None, k=1 << 17, R=4),
]


def data(n, p):
    X = qmc.Halton(p, scramble=False).random(n + 1)[1:]
    y = sum(np.sin(2 * np.pi * (i + 1) * X[:, i]) / (i + 1) for i in range(p))
    return np.ascontiguousarray(X), y


for case in CASES:
    n, p = case['n'], case['p']
    X, y = data(n, p)
    if case['kind'] == 'rbf':
        ls = 0.3 * (1 + np.arange(p) / 20.0)
        kern = ConstantKernel(1.0, 'fixed') * RBF(length_scale=ls, length_scale_bounds='fixed')
    else:
        kern = ConstantKernel(1.5, 'fixed') * Matern(length_scale=[case['ls']] * p, length_scale_bounds='fixed', nu=1.5)
    t0 = time.perf_counter()
    k = Kriging(X, y[:, None], kernel=kern)
    model = k._model()
    model.ensure_w()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    pts = torch.rand((case['k'], p), dtype=torch.float64, device='cuda')
    out = {}
    for want_std in (True, False):
        model.predict_device(pts[:4096], want_std)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            mean, std = model.predict_device(pts, want_std)
        e1.record()
        torch.cuda.synchronize()
        out['mean+sigma' if want_std else 'mean'] = 3 * case['k'] / (e0.elapsed_time(e1) * 1e-3)
    # parity spot check on the host (scikit-learn at the same theta)
    sub = pts[:256].cpu().numpy()
    ref = GaussianProcessRegressor(kernel=kern, normalize_y=True, optimizer=None).fit(X, y)
    rm, rs = ref.predict(sub, return_std=True)
    mean, std = model.predict_device(pts[:256], True)
    dm = float(np.max(np.abs(mean.cpu().numpy()[:, 0] - rm)) / max(ref._y_train_std, np.abs(rm).max()))
    dv = float(np.max(np.abs(std.cpu().numpy()[:, 0] ** 2 - rs ** 2)) / (ref._y_train_std ** 2 * 1.5))
    # batched LML
    yn, _, _ = normalise_y(y)
    Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda()
    free = ConstantKernel(1.5) * (RBF(length_scale=[1.0] * p) if case['kind'] == 'rbf' else Matern(length_scale=[1.0] * p, nu=1.5))
    rng = np.random.RandomState(0)
    specs = [kernel_spec.flatten(free.clone_with_theta(free.theta + rng.uniform(-0.5, 0.5, p + 1)), p) for _ in range(case['R'])]
    lml_batched(Xd, yd, specs[:2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vals = lml_batched(Xd, yd, specs)
    lml_s = time.perf_counter() - t0
    algo = n * n + 2 * n
    print(json.dumps(dict(case=case['name'], n=n, p=p, build_s=build_s, pred_per_s=out,
                          var_algorithmic_tflops=out['mean+sigma'] * algo * 1e-12,
                          mean_pairs_per_s=out['mean'] * n, parity_mean_rel=dm, parity_var_rel=dv,
                          lml_batch=case['R'], lml_batch_s=lml_s, lml_per_eval_ms=1e3 * lml_s / case['R'],
                          chol_tflops=case['R'] * (n ** 3 / 3.0) / lml_s * 1e-12, lml_finite=int(np.isfinite(vals).sum()))),
          flush=True)
    del k, model, pts
    torch.cuda.empty_cache()
