"""Where the config-3 step goes, from graph replays (warm, in-graph timing - ncu launch lists are cold and serialised):
python tools/perf_federer.py [envs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
BB = os.environ.get("BALL_BODY")          # A/B of vid2player.ball_body_contact: BALL_BODY=0 / 1
env = bench.federer_env(N, 0, **({"ball_body_contact": BB == "1"} if BB is not None else {}))
print("ball_body_contact:", env._physics_player.task._cfg_struct.ball_body_contact)
dev = env.device
task = env._physics_player.task
acts = [torch.clamp(torch.randn(N, env.num_actions, device=dev), -5, 5) for _ in range(8)]
for i in range(4):
    env.step(acts[i % 8]); env.reset(env.reset_buf.nonzero(as_tuple=False).flatten())
env.enable_cuda_graph()


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def graph_of(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


print(f"step + reset_done      : {timed(lambda: (env.step(acts[0]), env.reset_done())):.1f} us")
print(f"step graph alone       : {timed(lambda: env.step(acts[0])):.1f} us")
print(f"reset graph alone      : {timed(env.reset_done):.1f} us")
a75 = torch.zeros(N, task.num_actions, device=dev)
print(f"physics (3 launches)   : {timed(graph_of(lambda: task._env.step(a75))):.1f} us")
print(f"policy 4 layers        : {timed(graph_of(env._low_level_policy.forward_prepared)):.1f} us")
p = env._mvae_player
print(f"decoder + feedback     : {timed(graph_of(lambda: (p.decoder(env._mvae_actions), p.decoder.feed_back(3.0)))):.1f} us")
print(f"stream gather          : {timed(graph_of(lambda: p._gather(1))):.1f} us")
print(f"targets FK + obs       : {timed(graph_of(task.post_mvae_step)):.1f} us")
print(f"update_state           : {timed(graph_of(task._update_state_from_sim)):.1f} us")
print(f"controller post        : {timed(graph_of(lambda: env._compute_post(advance=True))):.1f} us")
print(f"empty graph replay     : {timed(graph_of(lambda: env._rng_done.zero_())):.1f} us")
