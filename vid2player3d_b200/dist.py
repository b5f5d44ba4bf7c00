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
    The bucket is allocated once and every parameter's `.grad` IS a view of it (set here, kept by `zero_grad(set_to_none=False)`
    or by calling `zero_()`): backward accumulates straight into the bucket, so a call is one collective + one scale, no packing.
    Parameters whose `.grad` was replaced (e.g. `zero_grad(set_to_none=True)`) are packed / unpacked with foreach copies instead.
    Every rank must hold the same trainable parameter set; the flat length is compared across ranks once at construction."""

    def __init__(self, params, group=None, grads_as_views=True):
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
        self.nbytes = n * 4
        if grads_as_views:
            for p, v in zip(self.params, self.views):
                p.grad = v
        if self.world > 1:
            lens = torch.tensor([n, len(self.params)], dtype=torch.int64, device=dev)
            lo, hi = lens.clone(), lens.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            if not (torch.equal(lo, lens) and torch.equal(hi, lens)):
                raise RuntimeError(f"GradAllReducer: ranks disagree on the trainable parameter set (this rank: {n} elements in "
                                   f"{len(self.params)} tensors; min {lo.tolist()}, max {hi.tolist()})")

    def zero_grad(self):
        """one memset of the bucket (parameters keep their views)"""
        self.flat.zero_()

    def synchronize(self):
        if self.world == 1:
            return
        stray = [(p, v) for p, v in zip(self.params, self.views) if p.grad is None or p.grad.data_ptr() != v.data_ptr()]
        if stray:
            torch._foreach_copy_([v for _, v in stray], [p.grad if p.grad is not None else torch.zeros_like(p) for p, _ in stray])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(self.world)
        for p, v in stray:
            if p.grad is None:
                p.grad = v
            else:
                p.grad.copy_(v)


def average_value(x, group=None):
    """hvd.average_value: mean of a scalar tensor over ranks"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    y = x.detach().clone().float()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y / dist.get_world_size(group)


def broadcast_parameters(model_or_params, src=0, group=None):
    """hvd.broadcast_parameters(model.state_dict()) at start-up (common_agent.py setup): pass the MODULE (or its state_dict) so
    that buffers - the RunningMeanStd statistics of the input / value normalisers - are broadcast too; an iterable of tensors /
    parameters is accepted for plain parameter lists."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    if isinstance(model_or_params, torch.nn.Module):
        tensors = list(model_or_params.state_dict().values())
    elif isinstance(model_or_params, dict):
        tensors = list(model_or_params.values())
    else:
        tensors = [p.data if isinstance(p, torch.nn.Parameter) else p for p in model_or_params]
    for t in tensors:
        if isinstance(t, torch.Tensor):
            dist.broadcast(t, src=src, group=group)


def broadcast_optimizer_state(optimizer, src=0, group=None):
    """hvd.broadcast_optimizer_state: tensor entries of the optimizer state (Adam moments, step counters) from rank `src`"""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    for st in optimizer.state.values():
        for k in sorted(st):
            if isinstance(st[k], torch.Tensor):
                dist.broadcast(st[k], src=src, group=group)
