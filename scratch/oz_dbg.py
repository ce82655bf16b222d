import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from batman_b200 import SurrogateModel
X, y = bench.training_set()
s = SurrogateModel('kriging', np.array([[0.0]*8, [1.0]*8]), None, kernel=bench.fixed_kernel(), global_optimizer=False)
s.fit(X, y)
m = s.predictor._model(); m.ensure_w()
pts = torch.rand((151552*4, 8), dtype=torch.float64, device='cuda')
from batman_b200 import _lib
import ctypes
lib = _lib.load()
for rep in range(2):
    lib.bk_profile_enable(1)
    m.predict_device(pts, True)
    torch.cuda.synchronize()
    ms = (ctypes.c_double*8)(); cnt = (ctypes.c_long*8)()
    lib.bk_profile_read(ms, cnt); lib.bk_profile_enable(0)
print("var ms/launch", ms[1]/cnt[1], "predict ms/launch", ms[0]/cnt[0])
