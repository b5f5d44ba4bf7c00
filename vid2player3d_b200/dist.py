"""Multi-GPU plumbing: one process per GPU (torchrun), envs sharded across ranks, ONE collective on the path -
the all-reduce of the flattened PPO gradient per minibatch.

Replaces the reference's optional Horovod call sites:
  hvd.rank()/seed offset                  embodied_pose/run.py:30-44
  optimizer.synchronize() (grad average)  learning/common_agent.py:388-395, agents/im_agent.py:553-560
  hvd.average_value(kl)                   learning/common_agent.py:180,196,202
  hvd.sync_stats                          learning/common_agent.py:95-96
The rollout itself needs no communication: envs are independent (collision group = env id).
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise from the torchrun environment; returns (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def shard_envs(total_envs, rank, world, pair=False):
    """Contiguous block of env ids for this rank.  Blocks are equal-sized; with pair=True (dual mode,
    opponent = id ^ 1, vid2player/utils/common.py:111-114) every block has even size so pairs never straddle GPUs."""
    if total_envs % world:
        raise ValueError(f"{total_envs} envs do not split evenly over {world} ranks")
    per = total_envs // world
    if pair and per % 2:
        raise ValueError("dual mode needs an even number of envs per rank")
    return range(rank * per, (rank + 1) * per)


def rank_seed(seed, rank):
    """cfg_train['params']['seed'] += rank  (run.py:37)"""
    return seed + rank


class GradAllReducer:
    """Averages the gradients of `params` across ranks with a single flat fp32 all-reduce (SUM, then / world).
    The bucket is allocated once; per call: pack -> all_reduce -> unpack (3 foreach ops + 1 collective)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def synchronize(self):
        if self.world == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(self.world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


def average_value(x, group=None):
    """hvd.average_value: mean of a scalar tensor over ranks"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    y = x.detach().clone().float()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y / dist.get_world_size(group)


def broadcast_parameters(params, src=0, group=None):
    """hvd.broadcast_parameters at start-up (setup_algo)"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in params:
            dist.broadcast(p.data, src=src, group=group)
