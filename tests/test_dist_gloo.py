"""world_size-2 gloo tests (CPU) of the N>1 host logic: env sharding and the PPO gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from vid2player3d_b200 import dist as D
    r, lr, w = D.init("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    torch.manual_seed(rank)        # different initial weights / buffers per rank: the broadcast must make them equal
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    net[1].running_mean.fill_(float(rank + 1))     # a buffer (stands for the RunningMeanStd statistics)
    D.broadcast_parameters(net)
    assert float(net[1].running_mean[0]) == 1.0, "buffers are broadcast with the parameters"
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    red = D.GradAllReducer(net.parameters())
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
    torch.manual_seed(100 + rank)  # rank-specific minibatch
    x, y = torch.randn(32, 8), torch.randn(32, 3)
    loss = ((net(x) - y) ** 2).mean()
    loss.backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.synchronize()
    synced = [p.grad.clone().numpy() for p in net.parameters()]
    opt.step()
    D.broadcast_optimizer_state(opt)
    red.zero_grad()
    assert all(float(p.grad.abs().max()) == 0.0 for p in net.parameters())
    kl = D.average_value(torch.tensor(float(rank + 1)))
    q.put((rank, [g.numpy() for g in local], synced, float(kl),
           list(D.shard_envs(16, rank, world, pair=True)), D.rank_seed(7, rank)))
    dist.destroy_process_group()


def test_grad_allreduce_and_sharding_world2():
    import numpy as np
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, l0, s0, kl0, ids0, seed0), (r1, l1, s1, kl1, ids1, seed1) = res
    for a, b, sa, sb in zip(l0, l1, s0, s1):
        np.testing.assert_allclose(sa, (a + b) / 2, rtol=1e-6, atol=1e-7)  # average of the two local grads
        np.testing.assert_allclose(sa, sb, rtol=0, atol=0)                 # identical on both ranks
    assert kl0 == kl1 == 1.5
    assert ids0 == list(range(0, 8)) and ids1 == list(range(8, 16))
    assert all((i ^ 1) in ids0 for i in ids0)  # dual-mode pairs stay on one rank
    assert (seed0, seed1) == (7, 8)


def test_shard_validation():
    from vid2player3d_b200 import dist as D
    with pytest.raises(ValueError):
        D.shard_envs(10, 0, 4)
    with pytest.raises(ValueError):
        D.shard_envs(6, 0, 2, pair=True)
    assert list(D.shard_envs(8192, 3, 8))[0] == 3072
