"""K5 timing breakdown: batched LML at n = 4096 / 1024 (run plain, then under ncu for the per-kernel list)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from sklearn.gaussian_process.kernels import ConstantKernel, Matern
from batman_b200 import kernel_spec
from batman_b200 import lml as L
from oracle import functions as ofun
n, p, R = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X = ofun.halton(n, p); y = ofun.g_function(X)[:, 0]
yn = (y - y.mean()) / y.std()
kern = ConstantKernel(1.0) * Matern(length_scale=[1.0] * p, nu=1.5)
rng = np.random.RandomState(0)
specs = [kernel_spec.flatten(kern.clone_with_theta(kern.theta + rng.uniform(-0.5, 0.5, p + 1)), p) for _ in range(R)]
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda()
L.lml_batched(Xd, yd, specs)
torch.cuda.synchronize()
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
t0 = time.perf_counter()
for _ in range(reps):
    out = L.lml_batched(Xd, yd, specs)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
flops = R * (n ** 3 / 3 + 2 * n * n)
print("n=%d p=%d R=%d: %.3f ms per batch, %.3f ms per LML, %.2f TFLOP/s algorithmic" % (n, p, R, dt * 1e3, dt * 1e3 / R, flops / dt * 1e-12))
