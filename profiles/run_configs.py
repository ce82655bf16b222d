"""BASELINE.json configs 1, 3, 4 through the public API (UQ.sobol), with wall times.  1 GPU, or N GPUs under
torchrun (rows shard, one all-reduce).  Profile script: may use oracle/ for data generation and checks.

    python profiles/run_configs.py [--cfg3-N 12500000] [--cfg4-N 10000000] > gpurun_out/configs.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from sklearn.gaussian_process.kernels import ConstantKernel, Matern  # noqa: E402

from batman_b200 import Kriging, Pod, SurrogateModel, UQ  # noqa: E402
from oracle import functions as ofun, gp as ogp, philox, pod as opod, saltelli  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--cfg3-N', type=int, default=12500000)
ap.add_argument('--cfg4-N', type=int, default=10000000)
args = ap.parse_args()
rank = int(os.environ.get('RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))


def timed_sobol(uq, reps=2):
    uq.sobol()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = uq.sobol()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return res, best


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


# ---- cfg1: Ishigami 3-D, n_train = 100 (the reference-fitted golden model), N = 1e4; full-size parity vs the oracle
from conftest import kriging_from_states, load_golden  # noqa: E402
z, states = load_golden('ishigami_n100')
s = SurrogateModel('kriging', z['corners'], None)
s.adopt(kriging_from_states(states), z['X_raw'])
N = 10000
dists = ['Uniform(-3.141592653589793, 3.141592653589793)'] * 3
(s2, s1, st), wall = timed_sobol(UQ(s, dists=dists, nsample=N, seed=123456))
A, B = philox.base_samples(123456, 0, N, [('uniform', -np.pi, np.pi)] * 3)
lo, hi = z['corners']
t0 = time.perf_counter()
Y = ogp.evaluate(states, (saltelli.design_from_base(A, B) - lo) / (hi - lo))[0]
want = saltelli.uq_sobol_aggregate(Y, N, 3)
cpu_s = time.perf_counter() - t0
a1, at, a2 = ofun.ishigami_indices()
emit(config="cfg1 Ishigami 3-D n_train=100 N=1e4", gpus=world, sobol_wall_s=wall, rows=8 * N,
     oracle_numpy_batched_wall_s=cpu_s, max_abs_diff_vs_oracle=float(max(np.abs(s1 - want[1]).max(), np.abs(st - want[2]).max(),
                                                                     np.abs(s2 - want[0]).max())),
     S1=s1.tolist(), ST=st.tolist(), analytic_S1=a1.tolist(), analytic_ST=at.tolist())

# ---- cfg3: Michalewicz 10-D, n_train = 4096, N per run given (1e8 / 8 GPUs = 1.25e7 per GPU)
p, n = 10, 4096
X = ofun.halton(n, p)
y = ofun.michalewicz(np.pi * X)
kern = ConstantKernel(1.0, 'fixed') * Matern(length_scale=[0.6] * p, length_scale_bounds='fixed', nu=1.5)
s3 = SurrogateModel('kriging', np.array([[0.0] * p, [np.pi] * p]), None, kernel=kern)
t0 = time.perf_counter()
s3.fit(np.pi * X, y)
fit_s = time.perf_counter() - t0
N3 = args.cfg3_N * world
(s2, s1, st), wall = timed_sobol(UQ(s3, dists=['Uniform(0., 3.141592653589793)'] * p, nsample=N3), reps=1)
emit(config="cfg3 Michalewicz 10-D n_train=4096", gpus=world, N=N3, rows=N3 * (2 * p + 2), sobol_wall_s=wall,
     mean_predictions_per_s=N3 * (2 * p + 2) / wall, pairs_per_s_per_gpu=N3 * (2 * p + 2) * n / wall / world,
     factorise_s=fit_s, S1=s1.tolist(), ST=st.tolist())

# ---- cfg4: Channel_Flow F = 400, p = 2, POD 20 modes + per-mode Kriging (n_train = 128), functional indices
corners = np.array([[15.0, 2500.0], [60.0, 6000.0]])
Xc = ofun.halton(128, 2, corners)
Yf = ofun.channel_flow(Xc, dx=100.0)
mean_snapshot, U, S, V = opod.fit(Yf.T, 1.0 - 1e-12, 20)
modes = V * S                                          # what the driver fits the surrogate on (pod.VS)
m = U.shape[1]
kern = ConstantKernel(1.0, 'fixed') * Matern(length_scale=[0.5, 0.5], length_scale_bounds='fixed', nu=1.5)
s4 = SurrogateModel('kriging', corners, None, kernel=kern)
s4.fit(Xc, modes, pod=Pod(mean_snapshot, U))
N4 = args.cfg4_N * world
uq4 = UQ(s4, dists=['Uniform(15., 60.)', 'Normal(4035., 400.)'], nsample=N4)
(s2, s1, st), wall = timed_sobol(uq4, reps=1)
emit(config="cfg4 Channel_Flow POD(%d modes) F=%d" % (m, Yf.shape[1]), gpus=world, N=N4, rows=N4 * 4, sobol_wall_s=wall,
     field_values_per_s=N4 * 4 * Yf.shape[1] / wall, S1_agg=s1.tolist(), ST_agg=st.tolist(),
     S1_map_first5=uq4.details['S1'][:5].tolist())
(b2, b1, bt), wallb = timed_sobol(UQ(s4, dists=['Uniform(15., 60.)', 'Normal(4035., 400.)'], nsample=N4, indices='block'), reps=1)
emit(config="cfg4 block indices", gpus=world, N=N4, sobol_wall_s=wallb, S1=b1.tolist(), ST=bt.tolist())
if world > 1:
    dist.destroy_process_group()
