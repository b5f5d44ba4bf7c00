"""Phase timing of the vid2player federer high-level step (not the bench contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import SIM_PARAMS, v2p_cfg
from vid2player3d_b200.tasks import PhysicsMVAEController
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(10)
env = PhysicsMVAEController(v2p_cfg(N), SIM_PARAMS, 1, "cuda", 0, True)
env.reset()
acts = [torch.clamp(torch.randn(N, 35, device=env.device), -5, 5) for _ in range(8)]
empty = torch.zeros(0, dtype=torch.long, device=env.device)
def timed(fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) / n * 1e3
for i in range(10): env.step(acts[i % 8]); env.reset(empty)
task = env._physics_player.task
print("phase: gpu-ms (event) / wall-ms")
print("pre_physics_step  ", timed(lambda i: env.pre_physics_step(acts[i % 8])))
print("  mvae_player.step", timed(lambda i: env._mvae_player.step(acts[i % 8][:, :32], acts[i % 8][:, 32:35])))
print("  post_mvae_step  ", timed(lambda i: task.post_mvae_step()))
print("physics_step      ", timed(lambda i: env.physics_step()))
print("  task.step       ", timed(lambda i: task.step(torch.zeros(N, 75, device=env.device))))
print("  native step only", timed(lambda i: task._env.step(torch.zeros(N, 75, device=env.device))))
print("post_physics_step ", timed(lambda i: env.post_physics_step()))
print("reset fast path   ", timed(lambda i: env.reset(empty)))
print("nonzero sync      ", timed(lambda i: env.reset_buf.nonzero().flatten()))
print("full step+reset   ", timed(lambda i: (env.step(acts[i % 8]), env.reset(env.reset_buf.nonzero().flatten()))))
env.enable_cuda_graph()
print("graph step only   ", timed(lambda i: env.step(acts[i % 8])))
print("graph step+reset  ", timed(lambda i: (env.step(acts[i % 8]), env.reset(env.reset_buf.nonzero().flatten()))))
print("resets pending:", int(env.reset_buf.sum()), "reaction", int(env._reset_reaction_buf.sum()))
ids = torch.tensor([3, 77, 500, 4000, 8000], device=env.device)
player = env._mvae_player
print("--- reset pieces with 5 humanoid ids")
print("player.reset      ", timed(lambda i: player.reset(ids)))
print("task._reset_actors", timed(lambda i: task._reset_actors(ids)))
print("  smpl_to_sim all ", timed(lambda i: task._smpl_to_sim_into(player._root_pos.contiguous(), player._joint_rotmat, task._tmp)))
print("tasks_fast(update)", timed(lambda i: env._reset_tasks_fast(update_state=True)))
print("  update_state    ", timed(lambda i: task._update_state_from_sim()))
print("_reset_envs(ids)  ", timed(lambda i: env._reset_envs(ids)))
print("nonzero           ", timed(lambda i: env.reset_buf.nonzero().flatten()))
