import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from batman_b200 import SurrogateModel, UQ
X, y = bench.training_set()
s = SurrogateModel('kriging', np.array([[0.0]*8, [1.0]*8]), None, kernel=bench.fixed_kernel(), global_optimizer=False)
s.fit(X, y)
uq = UQ(s, dists=['Uniform(0., 1.)']*8, nsample=1000000, seed=123456)
uq.sobol(); torch.cuda.synchronize()
ts = []
for _ in range(4):
    t0 = time.perf_counter(); r = uq.sobol(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
print("sobol N=1e6 wall", min(ts), "pairs/s", 1.8e7*1024/min(ts), "S1[0]", r[1][0])
