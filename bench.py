"""Benchmark of the Kriging-over-Saltelli hot path (BASELINE.json metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n-base 1000000] [--no-configs]

Workload (N = 1 GPU): BASELINE.json configs[1] -- G_Function 8-D, Kriging
n_train = 1024 (ConstantKernel x Matern-3/2 ARD at a fixed theta), Saltelli
N = 1e6 -> R = N(2p+2) = 1.8e7 design rows.  One *step* is one pass of the hot
path over the whole design: predictive mean AND sigma for every row (what the
reference's ``gp.predict(return_std=True)`` computes per row,
kriging.py:210-212).  ``value`` = rows / s with the design resident in HBM;
``e2e`` = the same through ``SurrogateModel.__call__`` with a plain (pageable)
numpy array in and numpy arrays out, host<->device copies in the timed region
(median of five calls after one untimed call, every call in ``e2e.runs_s``;
``e2e_pinned``: the same from a pinned buffer).  ``sobol`` = ``UQ.sobol()``
itself (fused mean-only path: the reference discards sigma in uq.py:201-202),
timed inside the clock-sampled region, all-reduce included.

N > 1 (torchrun, one rank per GPU): weak scaling, rank g owns base rows
[g N, (g+1) N) of an experiment of size G N; no data-path collective for
predictions, ONE all-reduce of the (hi, lo) sums for Sobol'.  ``sobol_strong``
is the strong-scaling point (the SAME N = 1e6 experiment split over G GPUs),
``sobol.max_abs_diff_vs_single_gpu`` the parity of the sharded indices with an
unsharded run of the same experiment on rank 0.

``roofline`` is quoted on the pipe the dominant kernel runs on: int8 tensor
ops of k_var_ozaki over the measured SUSTAINED ``tcgen05.mma kind::i8`` rate
(the kernel is timed inside a long step; the burst rate is reported too); the
FP64-equivalent rate is reported beside it.  ``roofline_sobol`` does the same
for the fused Saltelli kernel (FP64 pipe).  ``north_star`` and ``configs``
carry the other BASELINE configurations at their model sizes.

``--impl reference``: the reference CPU path on the host cores, rank 0 only:
the call sequence of batman's Kriging on scikit-learn (oracle/sk_reference.py;
batman itself cannot travel to the GPU box), BLAS on ALL host
cores also under torchrun (which exports OMP_NUM_THREADS=1).
"""
import os
import sys

if '--impl' in sys.argv and 'reference' in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core (set before numpy loads its BLAS)
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import ctypes  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import tempfile  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_DIM = 8
N_TRAIN = 1024
LENGTH_SCALES = [0.8, 1.2, 1.7, 2.2, 2.8, 3.4, 4.0, 4.6]
CONST = 2.0
SEED = 123456
# tcgen05.mma kind::i8, M128 N256 K32, SMEM operands, all 148 SMs (profiles/r02_int8_umma_probe.jsonl): best single
# launch (burst) and the same kernel launched back to back for 4 s (sustained: what a long step can reach under the
# board's power cap -- the convention MEASURED_PEAKS.json uses for bf16)
INT8_PEAK_TOPS_BURST = 4565.8
INT8_PEAK_TOPS = 4429.4


# ------------------------------------------------------------------------------------------------ synthetic workloads
def g_function(X):
    a = np.arange(1, X.shape[1] + 1, dtype=np.float64)
    f = np.ones(len(X))
    for i in range(X.shape[1]):
        f = f * ((np.abs(4.0 * X[:, i] - 2) + a[i]) / (1.0 + a[i]))
    return f[:, None]


def michalewicz(X, m=10):
    f = np.zeros(len(X))
    for i in range(X.shape[1]):
        f = f + np.sin(X[:, i]) * np.sin((i + 1) * X[:, i] ** 2 / np.pi) ** (2 * m)
    return (-f)[:, None]


def channel_flow(X, dx=100.0, length=40000.0, width=500.0, slope=5e-4, hinit=10.0):
    """Backwater curve of functions/analytical.py:411-435, whole-sample numpy (F = length / dx values per point)."""
    ks, q = X[:, 0], X[:, 1]
    xs = np.arange(dx, length + 1, dx)
    dl = int(length // dx)
    hc = np.power(q ** 2 / (9.8 * width ** 2), 1.0 / 3.0)
    hn = np.power(q ** 2 / (slope * width ** 2 * ks ** 2), 3.0 / 10.0)
    h = hinit * np.ones((len(X), dl))
    for i in range(2, dl + 1):
        prev = h[:, dl - i + 1]
        h[:, dl - i] = prev - dx * slope * ((1 - np.power(prev / hn, -10.0 / 3.0)) / (1 - np.power(prev / hc, -3.0)))
    return -xs[None, :] * slope + h


def halton(n, p, lo=None, hi=None):
    from scipy.stats import qmc
    X = qmc.Halton(p, scramble=False).random(n + 1)[1:]
    if lo is not None:
        X = np.asarray(lo) + (np.asarray(hi) - np.asarray(lo)) * X
    return X


def training_set(n=N_TRAIN, p=P_DIM):
    X = halton(n, p)
    return X, g_function(X)


def fixed_kernel(ls=LENGTH_SCALES, const=CONST, nu=1.5):
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    return ConstantKernel(const, constant_value_bounds='fixed') * Matern(
        length_scale=list(ls), length_scale_bounds='fixed', nu=nu)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_median=float(np.median(pw)))
        return out


def measure_dgemm_peak(torch):
    """cuBLAS DGEMM, torch.matmul fp64 8192^3 best of 10: same recipe as MEASURED_PEAKS.json's bf16 entry
    (that file has no FP64 number; profiles/r01_fp64_peaks.jsonl holds the raw DFMA/DMMA pipe rates)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return 2 * n ** 3 / best * 1e-12


def ncu_traffic_bytes(kernel):
    """DRAM bytes of one launch of `kernel` from the newest committed ncu capture (None if no summary is committed)."""
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    prof = os.path.join(ROOT, 'profiles')
    try:
        dirs = sorted((d for d in os.listdir(prof) if d.startswith('r0') and '_ncu' in d), reverse=True)
    except OSError:
        return None, None
    for d in dirs:
        path = os.path.join(prof, d, kernel + '_summary.csv')
        if not os.path.exists(path):
            continue
        total = 0.0
        for line in open(path):
            f = line.strip().split(',')
            if f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                try:
                    total += float(f[2]) * scale[f[1]]
                except (KeyError, ValueError, IndexError):
                    pass
        if total:
            return total, os.path.join('profiles', d, kernel + '_summary.csv')
    return None, None


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(steps, warmup, sample_rows, per_point_rows, design_rows):
    """Reference CPU path on the host cores (oracle/sk_reference.py): -> dict for the JSON line."""
    from threadpoolctl import threadpool_info, threadpool_limits
    from oracle import sk_reference
    cores = os.cpu_count() or 1
    threadpool_limits(limits=cores)                     # lift a launcher-imposed OMP_NUM_THREADS=1 (torchrun)
    X, y = training_set()
    t0 = time.perf_counter()
    k = sk_reference.fixed_theta_kriging(X, y, [fixed_kernel()])
    fit_s = time.perf_counter() - t0
    pts = design_rows[:sample_rows]
    for _ in range(warmup):
        k.evaluate_batched(pts[:8192])
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        k.evaluate_batched(pts)
        times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    k.evaluate(pts[:per_point_rows])
    shipped = per_point_rows / (time.perf_counter() - t0)
    threads = max([i.get('num_threads', 1) for i in threadpool_info()] + [1])
    if cores > 1:
        assert threads > 1, "the CPU reference arm must run BLAS on all host cores (got %d thread)" % threads
    # UQ.sobol() on the CPU: mean-only batched predictions over the design + the numpy restatement of
    # the Saltelli estimator, both on bounded samples and extrapolated linearly to N = 1e6 (R = 1.8e7 rows).  As
    # shipped the reference predicts mean AND sigma one point at a time and discards sigma (uq.py:192-203).
    from oracle import saltelli as osal
    t0 = time.perf_counter()
    k.evaluate_batched(pts, return_std=False)
    mean_rate = sample_rows / (time.perf_counter() - t0)
    n_est = 20000
    y_est = np.random.RandomState(SEED).rand(n_est * (2 * P_DIM + 2), 1)
    t0 = time.perf_counter()
    osal.uq_sobol_aggregate(y_est, n_est, P_DIM)
    est_s_per_base_row = (time.perf_counter() - t0) / n_est
    n_full = 1000000
    rows_full = n_full * (2 * P_DIM + 2)
    sobol_cpu = {"N": n_full, "rows": rows_full, "mean_only_batched_pred_per_s": mean_rate,
                 "estimator_s_extrapolated": est_s_per_base_row * n_full,
                 "wall_s_extrapolated_batched": rows_full / mean_rate + est_s_per_base_row * n_full,
                 "wall_s_extrapolated_as_shipped": rows_full / shipped + est_s_per_base_row * n_full,
                 "sample": "%d rows mean-only, estimator on N = %d" % (sample_rows, n_est)}
    return dict(value=sample_rows / float(np.mean(times)), ms_per_step=1e3 * float(np.mean(times)),
                cores=cores, blas_threads=threads, as_shipped_value=shipped, fit_s=fit_s, sobol=sobol_cpu,
                sample="%d design rows per step, sklearn predict(return_std=True) in blocks of 8192 on all host cores; "
                       "as_shipped = kriging.py per-point multi_eval loop on %d rows" % (sample_rows, per_point_rows))


def host_design(n_base, seed=SEED):
    """Saltelli design rows on the host (the reference arm only needs *some* rows of the same distribution)."""
    rng = np.random.RandomState(seed)
    A, B = rng.rand(n_base, P_DIM), rng.rand(n_base, P_DIM)
    blocks = [A, B]
    for i in range(P_DIM):
        E = A.copy(); E[:, i] = B[:, i]; blocks.append(E)
    for i in range(P_DIM):
        C = B.copy(); C[:, i] = A[:, i]; blocks.append(C)
    return np.concatenate(blocks, axis=0)


_RESULT_OUT = None


def _claim_stdout():
    """Exactly ONE line goes to the real stdout (the JSON result).  Everything else that writes to file descriptor 1
    -- the "NCCL version ..." banner of the first collective, library chatter -- is rerouted to stderr."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def _emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


# ------------------------------------------------------------------------------------------------ helpers (GPU arm)
class Ctx:
    pass


def _barrier(cx):
    cx.torch.cuda.synchronize()
    if cx.world > 1:
        cx.dist.barrier()


def _max_over_ranks(cx, v):
    t = cx.torch.tensor([float(v)], dtype=cx.torch.float64, device=cx.dev)
    if cx.world > 1:
        cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
    return float(t.item())


def _event_time(cx, fn, reps=1):
    """CUDA-event time of fn() on the current stream, seconds (min over reps), max over ranks."""
    torch = cx.torch
    best = 1e30
    for _ in range(reps):
        _barrier(cx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return _max_over_ranks(cx, best)


def _wall_time(cx, fn, reps=3):
    """Wall time of a host-driven call that ends synchronised (UQ.sobol returns host arrays), min over reps, max over
    ranks; a barrier on both sides."""
    best, out = 1e30, None
    for _ in range(reps):
        _barrier(cx)
        t0 = time.perf_counter()
        out = fn()
        cx.torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return _max_over_ranks(cx, best), out


def _profile(cx, fn):
    """Per-category CUDA-event durations of the library's own launches during fn()."""
    cx.lib.bk_profile_enable(1)
    fn()
    cx.torch.cuda.synchronize()
    cx.lib.bk_profile_enable(0)
    ms, cnt = (ctypes.c_double * 8)(), (ctypes.c_long * 8)()
    cx._lib.check(cx.lib.bk_profile_read(ms, cnt))
    return list(ms), list(cnt)


def _fixed_theta_surrogate(cx, X_raw, y, corners, kernel, pod=None):
    s = cx.SurrogateModel('kriging', corners, None, kernel=kernel, global_optimizer=False)
    s.fit(X_raw, y, pod=pod)
    return s


# ------------------------------------------------------------------------------------------------ other BASELINE configs
def north_star_block(cx):
    """8-D, n_train = 4096 (the model the north_star target is quoted on): mean+sigma and mean-only pred/s."""
    torch = cx.torch
    X, y = training_set(4096, P_DIM)
    s = _fixed_theta_surrogate(cx, X, y, np.array([[0.0] * P_DIM, [1.0] * P_DIM]), fixed_kernel())
    model = s.predictor._model()
    model.ensure_w()
    k_std, k_mean = 1 << 20, 1 << 22
    pts = torch.rand((k_mean, P_DIM), dtype=torch.float64, device=cx.dev)
    model.predict_device(pts[:8192], True)
    t_std = _event_time(cx, lambda: model.predict_device(pts[:k_std], True, check=False), reps=2)
    t_mean = _event_time(cx, lambda: model.predict_device(pts, False), reps=2)
    n = 4096
    return {"model": "G_Function 8-D, n_train=4096, Const*Matern32 ARD fixed theta", "gpus": cx.world,
            "mean_sigma_pred_per_s": cx.world * k_std / t_std, "mean_only_pred_per_s": cx.world * k_mean / t_mean,
            "rows_per_gpu": {"mean_sigma": k_std, "mean_only": k_mean},
            "fp64_equiv_tflops_per_gpu": (n * n + 2 * n + n * (3 * P_DIM + 8)) * k_std / t_std * 1e-12,
            "target": "north_star asks >= 1e10 mean+variance pred/s on 8 GPUs: n^2 flop per prediction puts that "
                      "three orders of magnitude above any FP64-exact roofline (BASELINE.md section 3)"}


def cfg3_block(cx, n_base):
    """configs[2]: Michalewicz 10-D, n_train = 4096, Saltelli N = 1e8 over 8 GPUs (1.25e7 base rows per GPU), plus one
    lockstep differential-evolution generation of all 32 restarts (32 x 15 x 11 candidates), sharded over the ranks."""
    torch = cx.torch
    p, n = 10, 4096
    X = halton(n, p)
    y = michalewicz(np.pi * X)
    s = _fixed_theta_surrogate(cx, np.pi * X, y, np.array([[0.0] * p, [np.pi] * p]), fixed_kernel([0.6] * p, 1.0))
    N = n_base * cx.world
    uq = cx.UQ(s, dists=['Uniform(0., 3.141592653589793)'] * p, nsample=N, seed=SEED)
    cx.UQ(s, dists=['Uniform(0., 3.141592653589793)'] * p, nsample=4096 * cx.world, seed=SEED).sobol()
    wall, ind = _wall_time(cx, uq.sobol, reps=1)
    rows = N * (2 * p + 2)
    out = {"config": "configs[2] Michalewicz 10-D n_train=4096, Saltelli N=%d (%d per GPU)" % (N, n_base), "gpus": cx.world,
           "rows": rows, "sobol_wall_s": wall, "mean_predictions_per_s": rows / wall,
           "fp64_fraction_of_dgemm_peak": rows / wall * n * (3 * p + 8) * 1e-12 / (cx.dgemm_peak * cx.world),
           "S1": [float(v) for v in ind[1]], "ST": [float(v) for v in ind[2]]}
    # 32 restarts, one generation each, in ONE lockstep batch (candidates shard over the ranks)
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    from batman_b200.fit import _Objective, normalise_y
    kern = ConstantKernel(1.0) * Matern(length_scale=[0.6] * p, nu=1.5)
    rng = np.random.RandomState(SEED)
    n_cand = 32 * 15 * (p + 1)
    lo, hi = np.log(0.05), np.log(20.0)
    thetas = np.column_stack([rng.uniform(-2, 2, n_cand)] + [rng.uniform(lo, hi, n_cand) for _ in range(p)])
    Xd = torch.from_numpy(np.ascontiguousarray(X)).to(cx.dev)
    Yd = torch.from_numpy(np.ascontiguousarray(normalise_y(y[:, 0])[0]))[None, :].to(cx.dev)
    obj = _Objective(kern, Xd, Yd, p)
    obj.neg_lml(thetas[:8 * cx.world], np.zeros(8 * cx.world))
    wall_fit, vals = _wall_time(cx, lambda: obj.neg_lml(thetas, np.zeros(n_cand)), reps=1)
    flop = n_cand * (n ** 3 / 3.0 + 2.0 * n ** 2)
    out["restarts"] = {"restarts": 32, "candidates_per_generation": n_cand, "n_train": n, "wall_s": wall_fit,
                       "lml_evaluations_per_s": n_cand / wall_fit, "algorithmic_tflops": flop / wall_fit * 1e-12,
                       "fraction_of_dgemm_peak": flop / wall_fit * 1e-12 / (cx.dgemm_peak * cx.world),
                       "finite": int(np.isfinite(vals).sum()),
                       "note": "one generation of all 32 restarts in lockstep = one batched call per rank"}
    return out


def cfg4_block(cx, n_base):
    """configs[3]: Channel_Flow field F = 400, POD with 20 modes + per-mode Kriging (n_train = 128), functional
    (aggregated + map) Sobol' indices, N = 1e7 base rows per GPU, p = 2 -> N(p+2) rows."""
    corners = np.array([[15.0, 2500.0], [60.0, 6000.0]])
    Xc = halton(128, 2, corners[0], corners[1])
    Yf = channel_flow(Xc)
    mean_snapshot = Yf.mean(axis=0)
    U, S, Vt = np.linalg.svd((Yf - mean_snapshot).T, full_matrices=False)      # pod.py:210-217
    U, modes = U[:, :20].copy(), (Vt.T * S)[:, :20].copy()
    s = _fixed_theta_surrogate(cx, Xc, modes, corners, fixed_kernel([0.5, 0.5], 1.0), pod=cx.Pod(mean_snapshot, U))
    dists = ['Uniform(15., 60.)', 'Normal(4035., 400.)']
    N = n_base * cx.world
    cx.UQ(s, dists=dists, nsample=4096 * cx.world, seed=SEED).sobol()
    uq = cx.UQ(s, dists=dists, nsample=N, seed=SEED)
    wall, ind = _wall_time(cx, uq.sobol, reps=1)
    return {"config": "configs[3] Channel_Flow F=%d, POD %d modes, n_train=128, Saltelli N=%d (%d per GPU)"
                      % (Yf.shape[1], U.shape[1], N, n_base), "gpus": cx.world, "rows": N * 4, "sobol_wall_s": wall,
            "field_values_per_s": N * 4 * Yf.shape[1] / wall, "S1_agg": [float(v) for v in ind[1]],
            "ST_agg": [float(v) for v in ind[2]]}


def cfg5_block(cx, k_per_gpu):
    """configs[4]: synthetic 20-D anisotropic RBF, n_train = 16384, mean+sigma prediction throughput."""
    torch = cx.torch
    p, n = 20, 16384
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel
    X = halton(n, p)
    y = sum(np.sin(2 * np.pi * (i + 1) * X[:, i]) / (i + 1) for i in range(p))[:, None]
    kern = ConstantKernel(1.0, 'fixed') * RBF(length_scale=0.3 * (1 + np.arange(p) / 20.0), length_scale_bounds='fixed')
    t0 = time.perf_counter()
    s = _fixed_theta_surrogate(cx, X, y, np.array([[0.0] * p, [1.0] * p]), kern)
    model = s.predictor._model()
    model.ensure_w()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    pts = torch.rand((k_per_gpu, p), dtype=torch.float64, device=cx.dev)
    model.predict_device(pts[:4096], True)
    t_std = _event_time(cx, lambda: model.predict_device(pts, True, check=False), reps=2)
    t_mean = _event_time(cx, lambda: model.predict_device(pts, False), reps=2)
    return {"config": "configs[4] synthetic 20-D RBF ARD n_train=16384, %d test points per GPU" % k_per_gpu,
            "gpus": cx.world, "mean_sigma_pred_per_s": cx.world * k_per_gpu / t_std,
            "mean_only_pred_per_s": cx.world * k_per_gpu / t_mean,
            "fp64_equiv_tflops_per_gpu": (n * n + 2 * n + n * (3 * p + 8)) * k_per_gpu / t_std * 1e-12,
            "sigma_kernel": "%s (scheme %d)" % (model.sigma_kernel, model.int8_scheme),
            "model_build_s": build_s}


# ------------------------------------------------------------------------------------------------ main
def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--n-base', type=int, default=1000000, help='Saltelli base size N per GPU')
    ap.add_argument('--cpu-sample', type=int, default=100000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the north_star / configs[2..4] blocks')
    ap.add_argument('--cfg3-n-base', type=int, default=12500000, help='configs[2] base rows per GPU (x8 = 1e8)')
    ap.add_argument('--cfg4-n-base', type=int, default=10000000, help='configs[3] base rows per GPU')
    ap.add_argument('--cfg5-points', type=int, default=1 << 18, help='configs[4] test points per GPU')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    nb = 2 * P_DIM + 2
    R = args.n_base * nb
    config = {"workload": "configs[1]: G_Function 8-D, Kriging n_train=1024 (ConstantKernel*Matern32 ARD, fixed theta), "
                          "Saltelli N=%d -> %d design rows per GPU, mean+sigma per row" % (args.n_base, R),
              "n_train": N_TRAIN, "p": P_DIM, "saltelli_N_per_gpu": args.n_base, "rows_per_gpu": R,
              "l2": "inputs larger than L2: %.2f GB design + K* chunks of 1.24 GB streamed per step" % (R * P_DIM * 8 / 1e9),
              "parallelism": "row-sharded x%d, no data-path collective (one all-reduce of Sobol' sums)" % world}
    metric, unit = "gp_mean_variance_predictions_per_sec", "pred/s"

    if args.impl == 'reference':
        if rank != 0:
            return
        sample = min(args.cpu_sample, R)
        ref = cpu_reference(args.steps, args.warmup, sample, 1000, host_design(max(2, sample // nb + 1)))
        line = {"impl": "reference", "metric": metric, "value": ref['value'], "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref['ms_per_step'],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": ref['value'], "unit": unit, "cores": ref['cores'], "kind": "port",
                                 "sample": ref['sample'], "blas_threads": ref['blas_threads'],
                                 "as_shipped_value": ref['as_shipped_value'], "sobol": ref['sobol']},
                "e2e": {"value": ref['value'], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(line)
        return

    lib_path = os.path.join(ROOT, 'batman_b200', 'libbatman_b200.so')
    if not os.path.exists(lib_path):                         # a checkout without the (git-ignored) built library
        if int(os.environ.get('LOCAL_RANK', '0')) == 0:
            from batman_b200.csrc import build as _build
            _build.build(force=True)
        else:
            deadline = time.time() + 900
            while not os.path.exists(lib_path) and time.time() < deadline:
                time.sleep(2.0)
            time.sleep(5.0)                                  # let the linker finish writing
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', local_rank))
    from batman_b200 import Pod, SurrogateModel, UQ, _lib
    lib = _lib.load()
    dev = torch.device('cuda', local_rank)
    cx = Ctx()
    cx.torch, cx.dist, cx.world, cx.rank, cx.dev, cx.lib, cx._lib = torch, dist, world, rank, dev, lib, _lib
    cx.SurrogateModel, cx.UQ, cx.Pod = SurrogateModel, UQ, Pod

    # ---- model (fit-at-fixed-theta on the device) and the resident design slab of this rank
    X, y = training_set()
    corners = np.array([[0.0] * P_DIM, [1.0] * P_DIM])
    surrogate = _fixed_theta_surrogate(cx, X, y, corners, fixed_kernel())
    model = surrogate.predictor._model()
    model.ensure_w()
    kinds = _lib.as_int_array([0] * P_DIM)
    params = _lib.as_double_array([0.0, 1.0, 0.0, 0.0] * P_DIM)        # Uniform(0, 1): 4 parameters per marginal
    row_lo = rank * args.n_base
    A = torch.empty((args.n_base, P_DIM), dtype=torch.float64, device=dev)
    B = torch.empty_like(A)
    _lib.check(lib.bk_design_rows(SEED, row_lo, args.n_base, P_DIM, kinds, params, 0, ctypes.c_void_p(A.data_ptr()), None))
    _lib.check(lib.bk_design_rows(SEED, row_lo, args.n_base, P_DIM, kinds, params, 1, ctypes.c_void_p(B.data_ptr()), None))
    D = UQ._design_blocks(A, B, nb).view(R, P_DIM)
    torch.cuda.synchronize()
    dists = ['Uniform(0., 1.)'] * P_DIM

    def step():
        return model.predict_device(D, return_std=True, check=False)

    uq = UQ(surrogate, dists=dists, nsample=args.n_base * world, seed=SEED)
    uq_ft = UQ(surrogate, dists=dists, nsample=args.n_base * world, seed=SEED, second_order=False)
    uq_strong = UQ(surrogate, dists=dists, nsample=args.n_base, seed=SEED)        # the SAME experiment at every G
    n_par = min(100000, args.n_base)
    uq_par = UQ(surrogate, dists=dists, nsample=n_par, seed=SEED)
    for _ in range(args.warmup):
        step()
    uq.sobol(); uq_ft.sobol(); uq_strong.sobol(); uq_par.sobol()
    model.check_pipeline_flag()
    _barrier(cx)

    # ================= timed, clock-sampled region: K prediction steps, then UQ.sobol() (all-reduce included)
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.bk_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mean, std = step()
    e1.record()
    _barrier(cx)
    lib.bk_profile_enable(0)
    ms_arr = (ctypes.c_double * 8)()
    cnt_arr = (ctypes.c_long * 8)()
    _lib.check(lib.bk_profile_read(ms_arr, cnt_arr))
    elapsed_ms = _max_over_ranks(cx, e0.elapsed_time(e1))
    value = world * R * args.steps / (elapsed_ms * 1e-3)
    model.check_pipeline_flag()
    sob_wall, indices = _wall_time(cx, uq.sobol, reps=3)
    sob_ft_wall, _ = _wall_time(cx, uq_ft.sobol, reps=3)
    strong_wall, ind_strong = _wall_time(cx, uq_strong.sobol, reps=3)
    clocks = sampler.stop()
    # ================= end of the timed region

    # sharded vs unsharded parity on the same experiment (N = n_par): every rank joins the sharded run, rank 0 repeats
    # it on its own GPU
    ind_par = uq_par.sobol()
    par_diff = 0.0
    if world > 1 and rank == 0:
        single = UQ(surrogate, dists=dists, nsample=n_par, seed=SEED, shard=False).sobol()
        par_diff = float(max(np.abs(a - b).max() for a, b in zip(ind_par, single)))
    sob_ms, sob_cnt = _profile(cx, uq.sobol)

    # ---- end to end through the public API: plain numpy in / numpy out, copies inside the timed region
    e2e_rows = R
    host_np = D.cpu().numpy()                                  # pageable, what a caller of SurrogateModel.__call__ holds
    surrogate(host_np[:4096])

    e2e_runs = {}

    def e2e_time(arr, tag):
        surrogate(arr)                                          # one untimed full-size call: staging buffers, streams
        times = []
        for _ in range(max(1, min(args.steps, 5))):
            _barrier(cx)
            t0 = time.perf_counter()
            surrogate(arr)
            times.append(time.perf_counter() - t0)
        e2e_runs[tag] = [round(t, 4) for t in times]
        # median of the calls, every call listed in `runs_s`: about one pageable call in four stalls for ~0.2 s on
        # these hosts, on every rank at once and whatever the staging scheme (profiles/r02_e2e_probe_4gpu.jsonl); the
        # same call fed from pinned memory never does
        return _max_over_ranks(cx, float(np.median(times)))
    e2e_value = world * e2e_rows / e2e_time(host_np, 'pageable')
    host_pin = torch.empty((e2e_rows, P_DIM), dtype=torch.float64, pin_memory=True)
    host_pin.copy_(D)
    e2e_pinned = world * e2e_rows / e2e_time(host_pin.numpy(), 'pinned')
    del host_pin

    cx.dgemm_peak = measure_dgemm_peak(torch)

    # ---- the other BASELINE configurations at their model sizes (collectives inside: every rank takes part)
    north_star = cfgs = None
    if not args.no_configs:
        del D, A, B, mean, std
        torch.cuda.empty_cache()
        north_star = north_star_block(cx)
        cfgs = []
        for fn, arg in ((cfg3_block, args.cfg3_n_base), (cfg4_block, args.cfg4_n_base), (cfg5_block, args.cfg5_points)):
            try:
                cfgs.append(fn(cx, arg))
            except Exception as exc:                 # informational blocks never break the bench line
                cfgs.append({"config": fn.__name__, "error": repr(exc)})
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- hyper-parameter fit side of the path, informational: one differential-evolution generation
    # (15 * (p + 1) = 135 candidates) of log-marginal-likelihood evaluations on this training set, one batched call
    fit_info = None
    try:
        from sklearn.gaussian_process.kernels import ConstantKernel, Matern
        from batman_b200 import kernel_spec
        from batman_b200.lml import lml_batched
        kern = ConstantKernel(CONST) * Matern(length_scale=list(LENGTH_SCALES), nu=1.5)
        rng = np.random.RandomState(SEED)
        n_cand = 15 * (P_DIM + 1)
        specs = [kernel_spec.flatten(kern.clone_with_theta(kern.theta + rng.uniform(-0.5, 0.5, P_DIM + 1)), P_DIM)
                 for _ in range(n_cand)]
        Xs_d = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
        yn = (y[:, 0] - y[:, 0].mean()) / y[:, 0].std()
        yn_d = torch.from_numpy(np.ascontiguousarray(yn)).to(dev)
        lml_batched(Xs_d, yn_d, specs)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            vals = lml_batched(Xs_d, yn_d, specs)
            best = min(best, time.perf_counter() - t0)
        flop = n_cand * (N_TRAIN ** 3 / 3.0 + 2.0 * N_TRAIN ** 2)
        fit_info = {"lml_evaluations_per_s": n_cand / best, "candidates_per_batch": n_cand, "n_train": N_TRAIN,
                    "ms_per_batch": best * 1e3, "finite": int(np.isfinite(vals).sum()),
                    "algorithmic_tflops": flop / best * 1e-12, "fraction_of_dgemm_peak": flop / best * 1e-12 / cx.dgemm_peak}
    except Exception as exc:                        # informational only: never break the bench line
        fit_info = {"error": repr(exc)}

    # ---- roofline of the dominant kernel, live CUDA-event durations on the launching stream
    var_ms, var_launches = ms_arr[1], cnt_arr[1]
    pred_ms, pred_launches = ms_arr[0], cnt_arr[0]
    algo_flops_per_pred = N_TRAIN ** 2 + 2 * N_TRAIN          # forward substitution + sum of squares
    rows_per_launch = R * args.steps / max(1, var_launches)
    avg_var_s = var_ms / max(1, var_launches) * 1e-3
    fp64_equiv = algo_flops_per_pred * rows_per_launch / avg_var_s * 1e-12
    int8 = model.sigma_kernel == 'int8'
    n_panels = (N_TRAIN + 127) // 128
    if int8:
        # executed int8 work of k_var_ozaki: 28 (7x8-bit scheme) or 36 (8x7-bit) slice pairs, each 128 test points x
        # 64 W rows x 32 k per slab; W is lower triangular, so the 64-row panel P runs 2 (P + 1) slabs: nP64 (nP64 + 1)
        # slabs per tile of 128 test points  (== ncu sm__ops_path_tensor_op_utcimma_src_int8)
        pairs = 28 if getattr(model, 'int8_scheme', 0) == 1 else 36
        n_p64 = 2 * n_panels
        int8_ops_per_row = pairs * 2.0 * 128 * 64 * 32 * n_p64 * (n_p64 + 1) / 128
        achieved = int8_ops_per_row * rows_per_launch / avg_var_s * 1e-12
        peak, r_unit = INT8_PEAK_TOPS, "TOP/s"
        kernel_name = "k_var_ozaki"
        kernel_desc = "k_var_ozaki (tcgen05 kind::i8 + TMEM: the FP64 product L^-1 k* evaluated exactly from int8 slices)"
        peak_source = ("tcgen05.mma kind::i8 of the best shape (M128 N256 K32) measured on this pool's B200s, sustained "
                       "figure (back-to-back launches for 4 s; the kernel is timed inside a long step), "
                       "profiles/r02_int8_umma_probe.jsonl; burst %.1f; MEASURED_PEAKS.json has bf16 and HBM entries only"
                       % INT8_PEAK_TOPS_BURST)
    else:
        achieved, peak, r_unit = fp64_equiv, cx.dgemm_peak, "TFLOP/s"
        kernel_name = "k_var_trmm"
        kernel_desc = "k_var_trmm (FP64 DMMA)"
        peak_source = "cuBLAS DGEMM measured live (torch.matmul fp64 8192^3, best of 10)"
    traffic, traffic_src = ncu_traffic_bytes(kernel_name)
    roofline = {"bound": "tensor", "kernel": kernel_desc, "achieved": achieved, "peak": peak, "unit": r_unit,
                "frac": achieved / peak, "frac_of_burst_peak": (achieved / INT8_PEAK_TOPS_BURST) if int8 else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": rows_per_launch * N_TRAIN * 8 + N_TRAIN * N_TRAIN * 8,
                "peak_source": peak_source,
                "mma_shapes": ("M128 K32, N = 64 x stacked W slices: 4 x N256, 2 x N192, 2 x N128, 2 x N64 per k slab "
                               "(7 x 8-bit scheme)") if int8 else None,
                "fp64_equiv": {"achieved_tflops": fp64_equiv, "dgemm_peak_tflops": cx.dgemm_peak,
                               "ratio_to_dgemm": fp64_equiv / cx.dgemm_peak,
                               "algorithmic_flops_per_prediction": algo_flops_per_pred},
                "avg_launch_ms": var_ms / max(1, var_launches), "launches": int(var_launches),
                "share_of_step": var_ms / elapsed_ms,
                "mean_kernel": {"kernel": "k_predict (FP64 pipe)", "avg_launch_ms": pred_ms / max(1, pred_launches),
                                "launches": int(pred_launches), "share_of_step": pred_ms / elapsed_ms,
                                "pairs_per_s": R * args.steps * N_TRAIN / (pred_ms * 1e-3),
                                "frac_of_dgemm_peak": R * args.steps * N_TRAIN * (3 * P_DIM + 8) / (pred_ms * 1e-3) * 1e-12 / cx.dgemm_peak}}
    sob_kernel_s = sob_ms[2] * 1e-3
    sob_rows = args.n_base * nb                                 # rows of this rank
    sob_flops = sob_rows * N_TRAIN * (3 * P_DIM + 8)            # n (3p + 8) per mean prediction
    roofline_sobol = {"bound": "fp64", "kernel": "k_sobol_static (design -> mean -> Saltelli sums, fused)",
                      "achieved": sob_flops / sob_kernel_s * 1e-12, "peak": cx.dgemm_peak, "unit": "TFLOP/s",
                      "frac": sob_flops / sob_kernel_s * 1e-12 / cx.dgemm_peak, "avg_launch_ms": sob_ms[2] / max(1, sob_cnt[2]),
                      "launches": int(sob_cnt[2]), "share_of_sobol_wall": sob_kernel_s / sob_wall,
                      "algorithmic_flops_per_prediction": N_TRAIN * (3 * P_DIM + 8),
                      "peak_source": "cuBLAS DGEMM measured live (the DFMA issue rate is 36.8 TFLOP/s, "
                                     "profiles/r01_fp64_peaks.jsonl); exp and sqrt count as one flop each, the SASS "
                                     "spends ~27 DP instructions on them"}

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(e2e_rows * P_DIM * 8),
                    "d2h_bytes_per_step": int(e2e_rows * 2 * 8),
                    "api": "SurrogateModel.__call__ (pageable numpy in, numpy out)", "runs_s": e2e_runs.get('pageable'),
                    "statistic": "median of the listed calls (after one untimed call), max over ranks"},
            "e2e_pinned": {"value": e2e_pinned, "unit": unit, "note": "same call fed from a pinned host buffer"},
            "gpu_launches": int(var_launches + pred_launches), "roofline": roofline, "roofline_sobol": roofline_sobol,
            "sobol": {"wall_s": sob_wall, "N": args.n_base * world, "rows": args.n_base * world * nb,
                      "mean_predictions_per_s": args.n_base * world * nb / sob_wall,
                      "first_total_only_wall_s": sob_ft_wall, "first_total_only_rows": args.n_base * world * (P_DIM + 2),
                      "timed_in_clock_sampled_region": True, "collective": "one all_reduce + one all_gather (intervals)",
                      "max_abs_diff_vs_single_gpu": par_diff, "parity_N": n_par,
                      "S1": [float(v) for v in indices[1]], "ST": [float(v) for v in indices[2]]},
            "sobol_strong": {"scaling": "strong", "N": args.n_base, "rows": args.n_base * nb, "gpus": world,
                             "wall_s": strong_wall, "S1": [float(v) for v in ind_strong[1]]},
            "fit": fit_info, "north_star": north_star, "configs": cfgs}
    if world == 1 and not args.no_cpu_baseline:
        ref = cpu_reference(2, 1, min(args.cpu_sample, R), 1000, host_np)
        line["cpu_baseline"] = {"value": ref['value'], "unit": unit, "cores": ref['cores'], "kind": "port",
                                "sample": ref['sample'], "blas_threads": ref['blas_threads'],
                                "as_shipped_value": ref['as_shipped_value'], "sobol": ref['sobol']}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
