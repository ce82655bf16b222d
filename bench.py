"""Benchmark of the Kriging-over-Saltelli hot path (BASELINE.json metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n-base 1000000]

Workload (N = 1 GPU): BASELINE.json configs[1] -- G_Function 8-D, Kriging
n_train = 1024 (ConstantKernel x Matern-3/2 ARD at a fixed theta), Saltelli
N = 1e6 -> R = N(2p+2) = 1.8e7 design rows.  One *step* is one pass of the hot
path over the whole design: predictive mean AND sigma for every row (what the
reference's ``gp.predict(return_std=True)`` computes per row,
kriging.py:210-212).  ``value`` = rows / s with the design resident in HBM;
``e2e`` = the same through ``SurrogateModel.__call__`` (numpy in / numpy out,
host<->device copies in the timed region).  ``sobol`` reports the wall time of
``UQ.sobol()`` itself (fused mean-only path: the reference discards sigma in
uq.py:201-202, so the drop-in never computes it there).

N > 1 (torchrun, one rank per GPU): weak scaling, rank g owns base rows
[g N, (g+1) N) of an experiment of size G N; no data-path collective for
predictions, ONE all-reduce of the (hi, lo) sums for Sobol'.

``--impl reference``: the reference CPU path on the host cores, rank 0 only:
the call sequence of batman's Kriging on scikit-learn (oracle/sk_reference.py;
batman itself cannot travel to the GPU box).  Its ``value`` is
the strongest fair CPU arm (same fitted model, block calls, BLAS on all cores);
the as-shipped per-point ``multi_eval`` loop is reported beside it.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_DIM = 8
N_TRAIN = 1024
LENGTH_SCALES = [0.8, 1.2, 1.7, 2.2, 2.8, 3.4, 4.0, 4.6]
CONST = 2.0
SEED = 123456


def g_function(X):
    a = np.arange(1, X.shape[1] + 1, dtype=np.float64)
    f = np.ones(len(X))
    for i in range(X.shape[1]):
        f = f * ((np.abs(4.0 * X[:, i] - 2) + a[i]) / (1.0 + a[i]))
    return f[:, None]


def training_set():
    from scipy.stats import qmc
    X = qmc.Halton(P_DIM, scramble=False).random(N_TRAIN + 1)[1:]
    return X, g_function(X)


def fixed_kernel():
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    return ConstantKernel(CONST, constant_value_bounds='fixed') * Matern(
        length_scale=LENGTH_SCALES, length_scale_bounds='fixed', nu=1.5)


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
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
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


def ncu_traffic_bytes(int8=False):
    """DRAM bytes of one launch of the dominant kernel from the committed ncu capture (None if the summary is absent)."""
    path = os.path.join(ROOT, 'profiles', 'r01_ncu_v5', 'k_var_ozaki_summary.csv') if int8 else \
        os.path.join(ROOT, 'profiles', 'r01_ncu_v2', 'k_var_trmm_summary.csv')
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    try:
        total = 0.0
        for line in open(path):
            f = line.strip().split(',')
            if f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                total += float(f[2]) * scale[f[1]]
        return total or None
    except Exception:
        return None


def cpu_reference(steps, warmup, sample_rows, per_point_rows, design_rows):
    """Reference CPU path on the host cores (oracle/sk_reference.py): -> dict for the JSON line."""
    from threadpoolctl import threadpool_info
    from oracle import sk_reference
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
                cores=os.cpu_count(), blas_threads=threads, as_shipped_value=shipped, fit_s=fit_s, sobol=sobol_cpu,
                sample="%d design rows per step, sklearn predict(return_std=True) in blocks of 8192 on all host cores; "
                       "as_shipped = kriging.py per-point multi_eval loop on %d rows" % (sample_rows, per_point_rows))


def host_design(n_base, seed=SEED):
    """Saltelli design rows on the host (numpy Philox is test infrastructure; the reference arm only needs
    *some* rows of the same distribution, so plain numpy uniforms are used)."""
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
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    nb = 2 * P_DIM + 2
    R = args.n_base * nb
    config = {"workload": "configs[1]: G_Function 8-D, Kriging n_train=1024 (ConstantKernel*Matern32 ARD, fixed theta), "
                          "Saltelli N=%d -> %d design rows per GPU, mean+sigma per row" % (args.n_base, R),
              "n_train": N_TRAIN, "p": P_DIM, "saltelli_N_per_gpu": args.n_base, "rows_per_gpu": R,
              "l2": "inputs larger than L2: %.2f GB design + K* chunks of 0.62 GB streamed per step" % (R * P_DIM * 8 / 1e9),
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
    from batman_b200 import Kriging, SurrogateModel, UQ, _lib
    lib = _lib.load()
    dev = torch.device('cuda', local_rank)

    # ---- model (fit-at-fixed-theta on the device) and the resident design slab of this rank
    X, y = training_set()
    corners = np.array([[0.0] * P_DIM, [1.0] * P_DIM])
    surrogate = SurrogateModel('kriging', corners, None, kernel=fixed_kernel(), global_optimizer=False)
    surrogate.fit(X, y)
    model = surrogate.predictor._model()
    model.ensure_w()
    kinds = _lib.as_int_array([0] * P_DIM)
    params = _lib.as_double_array([0.0, 1.0] * P_DIM)
    row_lo = rank * args.n_base
    A = torch.empty((args.n_base, P_DIM), dtype=torch.float64, device=dev)
    B = torch.empty_like(A)
    _lib.check(lib.bk_design_rows(SEED, row_lo, args.n_base, P_DIM, kinds, params, 0, ctypes.c_void_p(A.data_ptr()), None))
    _lib.check(lib.bk_design_rows(SEED, row_lo, args.n_base, P_DIM, kinds, params, 1, ctypes.c_void_p(B.data_ptr()), None))
    D = UQ._design_blocks(A, B, nb).view(R, P_DIM)
    torch.cuda.synchronize()

    def step():
        return model.predict_device(D, return_std=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.bk_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mean, std = step()
    e1.record()
    barrier()
    lib.bk_profile_enable(0)
    clocks = sampler.stop()
    ms_arr = (ctypes.c_double * 8)()
    cnt_arr = (ctypes.c_long * 8)()
    _lib.check(lib.bk_profile_read(ms_arr, cnt_arr))
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms.item())
    value = world * R * args.steps / (elapsed_ms * 1e-3)

    # ---- Sobol' wall time: UQ.sobol() on the same surrogate, N*world base rows sharded over the ranks
    uq = UQ(surrogate, dists=['Uniform(0., 1.)'] * P_DIM, nsample=args.n_base * world, seed=SEED)
    uq.sobol()
    sob_times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        indices = uq.sobol()
        torch.cuda.synchronize()
        sob_times.append(time.perf_counter() - t0)
    sob = torch.tensor([min(sob_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sob, op=dist.ReduceOp.MAX)
    # first and total order only (no C blocks: N(p+2) predictions); the reference always computes second order
    uq_ft = UQ(surrogate, dists=['Uniform(0., 1.)'] * P_DIM, nsample=args.n_base * world, seed=SEED, second_order=False)
    uq_ft.sobol()
    ft_times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        uq_ft.sobol()
        torch.cuda.synchronize()
        ft_times.append(time.perf_counter() - t0)
    sob_ft = torch.tensor([min(ft_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sob_ft, op=dist.ReduceOp.MAX)

    # ---- end to end through the public API with host buffers (pinned), copies inside the timed region
    e2e_rows = R
    host_pts = torch.empty((e2e_rows, P_DIM), dtype=torch.float64, pin_memory=True)
    host_pts.copy_(D[:e2e_rows])
    host_np = host_pts.numpy()
    surrogate(host_np[:4096])
    e2e_times = []
    for _ in range(max(1, min(args.steps, 3))):
        barrier()
        t0 = time.perf_counter()
        m_np, s_np = surrogate(host_np)
        e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_rows / float(e2e_t.item())

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
        fit_info = {"lml_evaluations_per_s": n_cand / best, "candidates_per_batch": n_cand, "n_train": N_TRAIN,
                    "ms_per_batch": best * 1e3, "finite": int(np.isfinite(vals).sum()),
                    "algorithmic_tflops": n_cand * (N_TRAIN ** 3 / 3.0 + 2.0 * N_TRAIN ** 2) / best * 1e-12}
    except Exception as exc:                        # informational only: never break the bench line
        fit_info = {"error": repr(exc)}

    # ---- roofline of the dominant kernel (gp_var_trmm), live CUDA-event durations on the launching stream
    var_ms, var_launches = ms_arr[1], cnt_arr[1]
    pred_ms, pred_launches = ms_arr[0], cnt_arr[0]
    algo_flops_per_pred = N_TRAIN ** 2 + 2 * N_TRAIN          # forward substitution + sum of squares
    rows_per_launch = R * args.steps / max(1, var_launches)
    achieved = algo_flops_per_pred * rows_per_launch / (var_ms / max(1, var_launches) * 1e-3) * 1e-12
    peak = measure_dgemm_peak(torch)
    int8 = model.sigma_kernel == 'int8'
    n_panels = (N_TRAIN + 127) // 128
    # executed int8 work of k_var_ozaki: 28 (7x8-bit scheme) or 36 (8x7-bit) slice-pair MMAs (M128 N128 K32) per 32-wide
    # k slab, 2 nP (nP+1) slabs per tile of 128 rows
    pairs = 28 if getattr(model, 'int8_scheme', 0) == 1 else 36
    int8_ops_per_row = pairs * 2.0 * 128 * 128 * 32 * 2 * n_panels * (n_panels + 1) / 128
    int8_tops = int8_ops_per_row * rows_per_launch / (var_ms / max(1, var_launches) * 1e-3) * 1e-12
    roofline = {"bound": "tensor",
                "kernel": "k_var_ozaki (tcgen05 int8 + TMEM: FP64 product evaluated exactly from int8 slices)" if int8
                          else "k_var_trmm (FP64 DMMA)",
                "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(int8),
                "note": ("frac > 1 is expected: achieved counts ALGORITHMIC FP64 flops (n^2 + 2n per prediction) while the "
                         "kernel runs them as 28 or 36 int8 MMAs per slab on the tcgen05 pipe; peak stays the measured FP64 (cuBLAS DGEMM) "
                         "rate so the number is comparable across rounds.  int8 pipe: see int8_tensor.") if int8 else None,
                "int8_tensor": {"achieved_tops": int8_tops, "peak_tops_same_shape": 3249.0, "peak_tops_n256": 4572.0,
                                "frac_same_shape": int8_tops / 3249.0,
                                "peak_source": "profiles/r01_int8_umma_probe.jsonl (tcgen05.mma kind::i8, SMEM operands)"} if int8 else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (151552 points) of the dominant "
                                  "kernel, ncu --set full capture in profiles/r01_ncu_v5 (int8) or r01_ncu_v2 (dmma); "
                                  "algorithmic bytes per launch = K* chunk + W = %.3e" % (151552 * N_TRAIN * 8 + N_TRAIN * N_TRAIN * 8),
                "peak_source": "measured live: torch.matmul fp64 8192^3 best of 10 (cuBLAS DGEMM); MEASURED_PEAKS.json "
                               "has no FP64 entry; raw DMMA pipe rate 37.1 TFLOP/s in profiles/r01_fp64_peaks.jsonl",
                "algorithmic_flops_per_prediction": algo_flops_per_pred,
                "avg_launch_ms": var_ms / max(1, var_launches), "launches": int(var_launches),
                "share_of_step": var_ms / elapsed_ms,
                "mean_kernel": {"kernel": "k_predict (FP64 pipe)", "avg_launch_ms": pred_ms / max(1, pred_launches),
                                "launches": int(pred_launches), "share_of_step": pred_ms / elapsed_ms,
                                "pairs_per_s": R * args.steps * N_TRAIN / (pred_ms * 1e-3)}}

    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(e2e_rows * P_DIM * 8),
                    "d2h_bytes_per_step": int(e2e_rows * 2 * 8), "api": "SurrogateModel.__call__ (numpy in, numpy out)"},
            "gpu_launches": int(var_launches + pred_launches), "roofline": roofline,
            "sobol": {"wall_s": float(sob.item()), "N": args.n_base * world, "rows": args.n_base * world * nb,
                      "mean_predictions_per_s": args.n_base * world * nb / float(sob.item()),
                      "first_total_only_wall_s": float(sob_ft.item()), "first_total_only_rows": args.n_base * world * (P_DIM + 2),
                      "S1": [float(v) for v in indices[1]], "ST": [float(v) for v in indices[2]]},
            "fit": fit_info}
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
