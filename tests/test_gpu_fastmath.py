"""GPU: the kernels' branch-free sqrt / exp against IEEE sqrt and a long-double exp."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(v):
    import torch
    from batman_b200 import _lib
    lib = _lib.load()
    x = torch.from_numpy(np.ascontiguousarray(v)).cuda()
    s, e = torch.empty_like(x), torch.empty_like(x)
    _lib.check(lib.bk_debug_fastmath(ctypes.c_void_p(x.data_ptr()), x.numel(), ctypes.c_void_p(s.data_ptr()),
                                     ctypes.c_void_p(e.data_ptr()), None))
    torch.cuda.synchronize()
    return s.cpu().numpy(), e.cpu().numpy()


def test_sqrt_is_correctly_rounded():
    """Distances must stay bit-identical to scipy's cdist (IEEE sqrt)."""
    rng = np.random.RandomState(0)
    v = np.concatenate([rng.rand(4000000) * 10.0 ** rng.uniform(-12, 6, 4000000),
                        np.array([0.0, 1.0, 4.0, 2.0, 1e-300, 1e300, 0.25, 3.0, np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0)]),
                        np.arange(1, 100001, dtype=np.float64) ** 2,
                        np.nextafter(np.arange(1, 100001, dtype=np.float64) ** 2, 0.0)])
    s, _ = _run(v)
    want = np.sqrt(v)
    bad = np.nonzero(s != want)[0]
    assert len(bad) == 0, (len(bad), v[bad[:5]], s[bad[:5]], want[bad[:5]])


def test_exp_neg_within_one_ulp():
    rng = np.random.RandomState(1)
    v = np.concatenate([rng.rand(1000000) * 50.0, rng.rand(200000) * 700.0, rng.rand(200000) * 1e-3,
                        np.array([0.0, 1e-300, 1e-17, 0.5, 1.0, np.log(2), 707.9, 708.1, 745.0, 800.0, 1e6])])
    _, e = _run(v)
    want = np.exp(-v.astype(np.longdouble))
    ulp = np.spacing(np.maximum(want.astype(np.float64), 1e-300))
    err = np.abs(e.astype(np.longdouble) - want) / ulp
    main = v < 708.0
    assert float(err[main].max()) <= 1.0, float(err[main].max())
    assert e[v == 0.0][0] == 1.0
    assert np.all(e[~main] == 0.0) and np.all(want[~main] < 3e-308)
    print("exp_neg: max %.3f ulp, mean %.3f ulp over %d values" % (err[main].max(), err[main].mean(), main.sum()))
