import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from batman_b200 import SurrogateModel
X, y = bench.training_set()
s = SurrogateModel('kriging', np.array([[0.0]*8, [1.0]*8]), None, kernel=bench.fixed_kernel(), global_optimizer=False)
s.fit(X, y)
R = 18000000
host = torch.empty((R, 8), dtype=torch.float64, pin_memory=True)
host.copy_(torch.rand((R, 8), dtype=torch.float64))
h = host.numpy()
s(h[:4096])
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); m, sg = s(h); best = min(best, time.perf_counter() - t0)
print("chunks", s.predictor._model().E2E_CHUNK // (148*128*8), "e2e pred/s %.4e  (%.1f ms)" % (R / best, best * 1e3))
