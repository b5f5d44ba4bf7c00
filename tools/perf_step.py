"""Quick A/B timing of the fused step kernel (not the bench contract): python tools/perf_step.py [envs] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import SIM_PARAMS, im_cfg
from vid2player3d_b200 import model_compiler, motion_lib
from vid2player3d_b200.tasks import HumanoidSMPLIM
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 320
model = model_compiler.load_compiled("smpl_mesh_humanoid_amass_v1")
flat = motion_lib.synthetic(model, num_motions=64, num_frames=300, seed=7)
torch.manual_seed(7)
task = HumanoidSMPLIM(im_cfg(N, flat), SIM_PARAMS, 1, "cuda", 0, True)
g = torch.Generator(device=task.device).manual_seed(1)
acts = [torch.rand(N, 75, device=task.device, generator=g) * 2 - 1 for _ in range(8)]
task.reset()
for i in range(20): task.step(acts[i % 8])
torch.cuda.synchronize()
tot = 0.0
for i in range(K):
    if i % 32 == 0: task.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); task.step(acts[i % 8]); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"{os.environ.get('B200ENV_LIB','default')} kernel={os.environ.get('B200ENV_KERNEL','packed')}: step_kernel avg {tot / K * 1e3:.1f} us -> {N * K / tot / 1e3:.2f} M env-steps/s "
      f"(rew mean {float(task.rew_buf.mean()):.4f}, resets {int(task.reset_buf.sum())})")
