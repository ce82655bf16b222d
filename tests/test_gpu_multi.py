"""GPU, 2 ranks (skipped on a single-GPU box): sharded UQ.sobol + NCCL all-reduce == single-GPU result."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    from conftest import kriging_from_states, load_golden
    from batman_b200 import SurrogateModel, UQ
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    z, states = load_golden('ishigami_n100')
    s = SurrogateModel('kriging', z['corners'], None)
    s.adopt(kriging_from_states(states), z['X_raw'])
    uq = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3, nsample=N)
    res = uq.sobol()
    if rank == 0:
        q.put([np.asarray(r) for r in res])
    dist.destroy_process_group()


def test_two_gpu_sobol_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from conftest import kriging_from_states, load_golden
    from batman_b200 import SurrogateModel, UQ
    N = 50001
    with socket.socket() as sck:
        sck.bind(('127.0.0.1', 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z, states = load_golden('ishigami_n100')
    s = SurrogateModel('kriging', z['corners'], None)
    s.adopt(kriging_from_states(states), z['X_raw'])
    want = UQ(s, dists=['Uniform(-3.141592653589793, 3.141592653589793)'] * 3, nsample=N).sobol()
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=0, atol=1e-12)


def _fit_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from batman_b200 import Kriging
    from oracle import functions as ofun
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    X = ofun.halton(48, 2)
    y = ofun.michalewicz(np.pi * X)
    k = Kriging(X, y)                                  # DE generations sharded over the ranks
    q.put((rank, np.asarray(k.hyperparameter[0]), k.gp_models[0].log_marginal_likelihood_value_))
    dist.destroy_process_group()


def test_two_gpu_sharded_restarts_agree_across_ranks():
    """Hyper-parameter candidates shard across ranks (all-gather of LML values): every rank must end on
    the same theta, and it must be at least as good as scikit-learn's multi-restart optimum."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    from oracle import functions as ofun
    with socket.socket() as sck:
        sck.bind(('127.0.0.1', 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(res[0][1], res[1][1])
    assert res[0][2] == res[1][2]
    X = ofun.halton(48, 2)
    y = ofun.michalewicz(np.pi * X)
    ref = GaussianProcessRegressor(kernel=ConstantKernel() * Matern(length_scale=(1.0, 1.0), length_scale_bounds=[(0.01, 100)] * 2),
                                   normalize_y=True, n_restarts_optimizer=20, random_state=0).fit(X, y[:, 0])
    assert res[0][2] >= ref.log_marginal_likelihood_value_ - 1e-3


def test_one_process_two_devices():
    """One process driving two GPUs: function attributes and the side stream are per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import kriging_from_states, load_golden
    from batman_b200 import Kriging
    from sklearn.gaussian_process import kernels as sk
    z, states = load_golden('kernel_default_3out')
    with torch.cuda.device(0):
        m0, s0 = kriging_from_states(states).evaluate(z['points'])
    with torch.cuda.device(1):
        m1, s1 = kriging_from_states(states).evaluate(z['points'])
        X = states[0]['X_train']
        y = np.sin(3 * X[:, :1]) + X[:, 1:2]
        kern = sk.ConstantKernel(1.0, 'fixed') * sk.Matern([0.5] * X.shape[1], 'fixed', nu=1.5)
        f1 = Kriging(X, y, kernel=kern).evaluate(z['points'])
    with torch.cuda.device(0):
        f0 = Kriging(X, y, kernel=kern).evaluate(z['points'])
    np.testing.assert_array_equal(m0, m1)
    np.testing.assert_array_equal(s0, s1)
    np.testing.assert_array_equal(f0[0], f1[0])
    np.testing.assert_array_equal(f0[1], f1[1])
