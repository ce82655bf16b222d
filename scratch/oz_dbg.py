import sys, time, numpy as np, torch, ctypes
sys.path.insert(0, '.')
import bench
from batman_b200 import SurrogateModel, _lib
X, y = bench.training_set()
s = SurrogateModel('kriging', np.array([[0.0]*8, [1.0]*8]), None, kernel=bench.fixed_kernel(), global_optimizer=False)
s.fit(X, y)
m = s.predictor._model(); m.ensure_w()
pts = torch.rand((151552*2, 8), dtype=torch.float64, device='cuda')
lib = _lib.load()
best = [1e9, 1e9]
for rep in range(6):
    lib.bk_profile_enable(1)
    m.predict_device(pts, True, check=False)
    torch.cuda.synchronize()
    ms = (ctypes.c_double*8)(); cnt = (ctypes.c_long*8)()
    lib.bk_profile_read(ms, cnt); lib.bk_profile_enable(0)
    best = [min(best[0], ms[1]/cnt[1]), min(best[1], ms[0]/cnt[0])]
print("scheme", m.int8_scheme, "var ms/launch %.3f predict ms/launch %.3f" % tuple(best))
