"""Which warp of the one-wave physics launch ends last, and in what phase?  Needs a -DPT_PROF=1 build (tools/pt_prof.sh):
per-warp cycle counts of the phases of the last control step, read back through b200env_debug_prof.
   B200ENV_LIB=vid2player3d_b200/lib/ab_prof.so python tools/pt_prof.py [rollout steps ...]"""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from vid2player3d_b200 import native

N = 8192
AMASS = len(sys.argv) > 1 and sys.argv[1] == "amass"      # config 2 (embodied_pose task, random policy) instead of the primary workload
steps = [int(a) for a in sys.argv[(2 if AMASS else 1):]] or ([5, 30] if AMASS else [10, 45, 55, 70])
if AMASS:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import SIM_PARAMS, im_cfg
    from vid2player3d_b200 import model_compiler, motion_lib
    from vid2player3d_b200.tasks import HumanoidSMPLIM
    model = model_compiler.load_compiled("smpl_mesh_humanoid_amass_v1")
    flat = motion_lib.synthetic(model, num_motions=64, num_frames=300, seed=7)
    torch.manual_seed(7)
    task = HumanoidSMPLIM(im_cfg(N, flat), SIM_PARAMS, 1, "cuda", 0, True)
    g = torch.Generator(device=task.device).manual_seed(1)
    acts = [torch.rand(N, 75, device=task.device, generator=g) * 2 - 1 for _ in range(8)]
    task.reset()

    class _E:      # the two calls the loop below makes
        def step(self, a): task.step(a)
        def reset_done(self): pass
    env = _E()
else:
    env = bench.federer_env(N, 0)
    dev = env.device
    acts = [torch.clamp(torch.randn(N, env.num_actions, device=dev), -5, 5) for _ in range(8)]
    for i in range(4):
        env.step(acts[i]); env.reset(env.reset_buf.nonzero(as_tuple=False).flatten())
    env.enable_cuda_graph()
names = ["barrier", "body pass", "contact", "backward", "root+ball", "last fwd", "-", "forward"]
buf = np.zeros((4096, 8), np.uint64)
nw = (N + 55) // 56 * 14
for i in range(max(steps) + 1):
    env.step(acts[i % 8]); env.reset_done()
    if i in steps:
        torch.cuda.synchronize()
        assert native.lib().b200env_debug_prof(buf.ctypes.data_as(C.c_void_p), C.c_int(buf.size)) == 0
        t = buf[:nw].astype(np.float64)
        t = t[t.sum(1) > 0]
        tot = t.sum(1)
        slow = np.argsort(tot)[-8:]
        print(f"rollout step {i}: {len(t)} warps, control step cycles mean {tot.mean():.0f}  p99 {np.percentile(tot, 99):.0f}  max {tot.max():.0f}  (max / mean {tot.max() / tot.mean():.2f})")
        print("   phase          mean      p99      max | mean over the 8 slowest warps")
        for k in (0, 1, 2, 3, 4, 7, 5):
            print(f"   {names[k]:10s} {t[:, k].mean():9.0f} {np.percentile(t[:, k], 99):8.0f} {t[:, k].max():8.0f} | {t[slow, k].mean():9.0f}")
