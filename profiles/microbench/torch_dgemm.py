"""cuBLAS DGEMM peak, measured the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3, best of 10)."""
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e-3)
print(json.dumps({"kernel": "torch.matmul fp64 8192^3 (cuBLAS DGEMM)", "tflops": 2 * n**3 / best * 1e-12}))
import time
t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    c = a @ b; k += 1
    if k % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"kernel": "torch.matmul fp64 8192^3 sustained 4s", "tflops": 2 * n**3 * k / (e0.elapsed_time(e1) * 1e-3) * 1e-12}))
