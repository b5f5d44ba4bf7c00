"""Kernel timeline of the primary workload's step (step graph + reset graph) from CUPTI (torch.profiler): start, duration and the idle gap in
front of every kernel - where the step's time is NOT inside a kernel.   python tools/step_timeline.py [envs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = bench.federer_env(N, 0)
dev = env.device
acts = [torch.clamp(torch.randn(N, env.num_actions, device=dev), -5, 5) for _ in range(4)]
for i in range(4):
    env.step(acts[i]); env.reset(env.reset_buf.nonzero(as_tuple=False).flatten())
env.enable_cuda_graph()
for i in range(10):
    env.step(acts[i % 4]); env.reset_done()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6):
        env.step(acts[i % 4]); env.reset_done()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ev.sort(key=lambda e: e.time_range.start)
phys = [i for i, e in enumerate(ev) if e.name.startswith("step_kernel") or "step_kernel_packed" in e.name]
if len(phys) < 4:
    print("no per-kernel records for graph replays:", len(ev), "events"); sys.exit(0)
a, b = phys[2], phys[3]          # one full period: physics launch of step t .. physics launch of step t + 1
tot_k = tot_g = 0.0
prev_end = None
for e in ev[a:b]:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = 0.0 if prev_end is None else s - prev_end
    prev_end = max(prev_end or 0, e.time_range.end)
    tot_k += d; tot_g += max(gap, 0.0)
    print(f"{gap:7.1f} us gap  {d:8.1f} us  {e.name[:100]}")
print(f"period {ev[b].time_range.start - ev[a].time_range.start:.1f} us: in kernels {tot_k:.1f} us, gaps {tot_g:.1f} us over {b - a} launches")
