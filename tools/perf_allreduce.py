"""The ONE collective of the training path (BASELINE config 4): GradAllReducer.synchronize() on an actor-critic of the reference's
size (~4.69 M parameters = 18.8 MB of fp32 gradients per minibatch), timed with CUDA events, max over ranks.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/perf_allreduce.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from vid2player3d_b200 import dist as D

rank, local_rank, world = D.init()
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
torch.manual_seed(D.rank_seed(0, rank))
# the embodied_pose actor-critic: 734-d obs -> [2048, 1024, 512] -> 75 (+ value head); sizes from cfg/*/train yaml (mlp units)
dims = [734, 2048, 1024, 512]
layers = []
for a, b in zip(dims[:-1], dims[1:]):
    layers += [torch.nn.Linear(a, b), torch.nn.ReLU()]
net = torch.nn.Sequential(*layers, torch.nn.Linear(dims[-1], 75 + 1)).to(dev)
D.broadcast_parameters(net.parameters())
red = D.GradAllReducer(net.parameters())
x = torch.randn(512, 734, device=dev)
net(x).sum().backward()
for _ in range(5):
    red.synchronize()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
ms = []
for _ in range(50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); red.synchronize(); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
t = torch.tensor([sorted(ms)[len(ms) // 2]], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
n = red.flat.numel()
# check: every rank ends with the same averaged gradient
chk = red.flat.double().sum().reshape(1)
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(hi - lo) == 0.0
if rank == 0:
    b = n * 4
    print(json.dumps({"world": world, "params": n, "bytes": b, "median_ms": float(t), "algbw_GBps": b / 1e9 / (float(t) * 1e-3),
                      "busbw_GBps": b / 1e9 / (float(t) * 1e-3) * 2 * (world - 1) / max(world, 1)}))
if world > 1:
    dist.destroy_process_group()
