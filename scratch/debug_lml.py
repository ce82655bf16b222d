import numpy as np, torch, sys
sys.path.insert(0, '.')
from scipy.linalg import cholesky, solve_triangular
from sklearn.gaussian_process.kernels import ConstantKernel, Matern
from batman_b200 import kernel_spec, _lib
from batman_b200 import lml as L
from batman_b200.fit import normalise_y
from oracle import functions as ofun, gp as ogp
n, p = 60, 3
X = ofun.halton(n, p); y = ofun.g_function(X)[:, 0]
yn, _, _ = normalise_y(y)
kern = ConstantKernel(1.0) * Matern(length_scale=[1.0]*p, nu=1.5)
spec = kernel_spec.flatten(kern, p)
print(spec)
Xd, yd = torch.from_numpy(X).cuda(), torch.from_numpy(yn).cuda()
got = L.lml_batched(Xd, yd, [spec])
print("lml got", got, "want", ogp.lml(ogp.from_sklearn(kern), X, yn))
ws = L._ws_cache[Xd.device]
raw = ws.cpu().numpy()
def al(v): return (v + 255)//256*256
n_pad = 128; nP = 1; R = 1
off = al(288*R)
M = raw[off:off + n_pad*n_pad*8].view(np.float64).reshape(n_pad, n_pad); off += al(n_pad*n_pad*8*R)
Winv = raw[off:off + nP*128*128*8].view(np.float64).reshape(128,128); off += al(nP*128*128*8*R)
z = raw[off:off+n_pad*8].view(np.float64); off += al(n_pad*8*R)
acc = raw[off:off+16].view(np.float64); off += al(16*R)
info = raw[off:off+4].view(np.int32); off += al(4*R)
ypad = raw[off:off+n_pad*8].view(np.float64)
K = ogp.kernel_train(ogp.from_sklearn(kern), X); K[np.diag_indices_from(K)] += 1e-10
Lr = cholesky(K, lower=True)
print("L err", np.abs(np.tril(M[:n,:n]) - Lr).max(), "pad diag", M[n:, n:].diagonal()[:3])
print("ypad err", np.abs(ypad[:n]-yn).max(), ypad[n:n+3])
zr = solve_triangular(Lr, yn, lower=True)
print("z err", np.abs(z[:n]-zr).max(), "z pad", np.abs(z[n:]).max())
print("Winv err", np.abs(Winv[:n,:n]-np.linalg.inv(Lr)).max())
print("acc", acc, "want logdet", np.log(np.diag(Lr)).sum(), "zz", zr@zr, "info", info)
Lm = np.tril(M[:n,:n]); Kp = Lm@Lm.T
print("recon K' vs K max err", np.abs(Kp-K).max())
print("K'[:4,:4]\n", Kp[:4,:4], "\nK[:4,:4]\n", K[:4,:4])
print("diag K'", Kp.diagonal()[:6])
