"""The physics launch of the primary workload k steps after a synchronous reset of all envs (every humanoid hits the ground ~1.5 - 2 s in):
python tools/transient_probe.py [steps_before] - the cudaProfiler range covers ONE step, for `ncu --profile-from-start off`; without ncu it
prints the step time at that point."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

K = int(sys.argv[1]) if len(sys.argv) > 1 else 55
N = 8192
BB = os.environ.get("BALL_BODY")          # A/B of vid2player.ball_body_contact: BALL_BODY=0 / 1
env = bench.federer_env(N, 0, **({"ball_body_contact": BB == "1"} if BB is not None else {}))
dev = env.device
acts = [torch.clamp(torch.randn(N, env.num_actions, device=dev), -5, 5) for _ in range(8)]
for i in range(4):
    env.step(acts[i]); env.reset(env.reset_buf.nonzero(as_tuple=False).flatten())
env.enable_cuda_graph()
ms = []
for i in range(K + 12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if i == K:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    e0.record(); env.step(acts[i % 8]); env.reset_done(); e1.record()
    if i == K:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
task = env._physics_player.task
z = task._rigid_body_state.view(N, -1, 13)[:, 0, 2]
print("kernel form", task._env.kernel_form, "step ms by rollout step (every 5th):", " ".join(f"{i}:{ms[i]:.3f}" for i in range(0, len(ms), 5)))
print(f"pelvis height < 0.5 m: {float((z < 0.5).float().mean()):.2f} of the envs; bodies with a ground-contact force: "
      f"{float((task._contact_forces.view(N, -1, 3)[:, :, 2].abs() > 1).float().sum(1).mean()):.1f} per env")
