#!/usr/bin/env python
"""bench.py - env-steps/s of the vid2player3d rollout hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs 8192] [--impl b200|reference]

One "step" = one `task.step(actions)` over all envs of a rank (2 sim steps x 2 substeps of the articulated
step + MoCap target + obs + reward + reset), random policy, with `task.reset()` of all envs every 32 steps
like the reference's ImitatorAgent.play_steps (agents/im_agent.py:305-409) - resets are inside the timed
region.  Workload = embodied_pose amass_im (BASELINE config 2 at the metric's 8192 envs): SMPL humanoid,
24 bodies / 69 dof, synthetic MoCap library (64 motions x 300 frames, seed 7), episodeLength 300.
Multi-GPU: envs shard across ranks, no data-path collective ("weak": 8192 envs per GPU).

Timing: every step is bracketed by CUDA events on the launching stream; an L2 flush (256 MiB memset) runs
between timed steps, outside the event pairs.  value = envs * K / sum(step times), max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALGO_BYTES_PER_ENV_STEP = 9792  # SURVEY.md 8(d), embodied_pose configs; derivation in DESIGN.md 5
# dram__bytes_read.sum + dram__bytes_write.sum of ONE step_kernel_packed<split> launch at 8192 envs, from the `ncu --set full`
# capture profiles/r1h_step_kernel_packed_ncu.md (7.51 MB read + 512 B written: the state rows the neighbouring launches of a
# rollout touch stay in the 126 MB L2, so the DRAM traffic is far BELOW the 80.2 MB of algorithmic bytes; pre_kernel adds 18.6 MB,
# post_kernel 40.3 MB, profiles/r1g ncu).  A constant from the profile, not measured by bench.py.
NCU_TRAFFIC_BYTES_PER_LAUNCH = 7508992
# FP32 work of one env step (SURVEY.md 8d asks for achieved FP32 FLOP/s next to the HBM figure): from the executed-instruction mix of
# the physics launch in the ncu source page of capture G (97.4 M warp instructions x 13.4 active lanes; 25.4 % FFMA = 2 flop, 18.0 % FMUL,
# 13.4 % FADD) = 1.07 GFLOP per 8192-env launch = 131 kFLOP per env step (pre / post launches add ~4 %); a constant from the profile.
FP32_FLOP_PER_ENV_STEP = 131.0e3
HORIZON = 32


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=1600)
    p.add_argument("--warmup", type=int, default=64)
    p.add_argument("--envs", type=int, default=8192, help="envs per GPU")
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--cpu-sample-envs", type=int, default=1024)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-federer", action="store_true", help="skip the secondary vid2player federer (config 3) measurement")
    return p.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(envs, device_index, seed):
    import torch
    from helpers import SIM_PARAMS, im_cfg
    from vid2player3d_b200 import model_compiler, motion_lib
    from vid2player3d_b200.tasks import HumanoidSMPLIM, VecTaskPythonWrapper
    model = model_compiler.load_compiled("smpl_mesh_humanoid_amass_v1")
    flat = motion_lib.synthetic(model, num_motions=64, num_frames=300, seed=7, sigma=0.05)
    torch.manual_seed(seed)
    task = HumanoidSMPLIM(im_cfg(envs, flat, episodeLength=300), SIM_PARAMS, 1, "cuda", device_index, True)
    return model, flat, task, VecTaskPythonWrapper(task, task.device, 5.0, 1.0)


def federer_workload(envs, device_index, steps=96, warmup=16):
    """BASELINE config 3 (vid2player federer single-player): 8192 envs, humanoid + racket + ball, substeps 6, episode 300,
    reward return_w_estimate with synthetic estimator tables, synthetic motion generator in place of the (unreleased) MVAE and a
    zero-residual low-level policy.  Returns high-level env-steps/s and the per-launch split."""
    import torch
    from helpers import SIM_PARAMS, v2p_cfg
    from vid2player3d_b200.tasks import PhysicsMVAEController
    torch.manual_seed(10)
    cfg = v2p_cfg(envs)
    cfg["env"]["motion_player"] = "stream"      # resident kinematic target stream in place of the unreleased MVAE
    env = PhysicsMVAEController(cfg, SIM_PARAMS, 1, "cuda", device_index, True)
    dev = env.device
    env.reset()
    acts = [torch.clamp(torch.randn(envs, env.num_actions, device=dev), -5, 5) for _ in range(8)]

    def run(n):
        for i in range(n):
            env.step(acts[i % 8])
            env.reset(env.reset_buf.nonzero(as_tuple=False).flatten())
    run(warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1)
    # whole high-level step captured in a CUDA graph (resets stay eager between replays)
    env.enable_cuda_graph()
    run(warmup)
    torch.cuda.synchronize()
    l0 = env._physics_player.task._env.launch_count
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # physics launch alone
    task = env._physics_player.task
    a75 = torch.zeros(envs, task.num_actions, device=dev)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(20):
        task._env.step(a75)
    p1.record()
    torch.cuda.synchronize()
    return {"env_steps_per_s": envs * steps / (ms * 1e-3), "ms_per_step": ms / steps, "physics_kernel_ms": p0.elapsed_time(p1) / 20,
            "mode": "one CUDA graph per high-level step + reset(done ids) = nonzero() host sync, mask fill, one reset-graph replay",
            "env_steps_per_s_eager": envs * steps / (ms_eager * 1e-3),
            "steps": steps, "step_kernel_launches_outside_graph": task._env.launch_count - l0 - 20,
            "workload": f"vid2player federer single: {envs} envs, humanoid+racket+ball, substeps 6 (12 per step), return_w_estimate, "
                        "resident kinematic target stream (48 frames in HBM) + zero-residual low-level policy (MVAE / policy checkpoints unreleased)",
            "roofline_frac_hbm": 10900 * envs / (p0.elapsed_time(p1) / 20 * 1e-3) / 1e9 / peaks()[0]}


def dual_workload(envs, device_index, steps=96, warmup=16):
    """BASELINE config 5 (vid2player federer_djokovic dual) on one GPU: `envs` paired envs (envs/2 rallies), two assets stepped by
    two launches per step (even rows = federer, odd rows = djokovic), dual reset FSM, ball hand-over through the incoming-ball
    table; synthetic motion generator / zero-residual low-level policy as in config 3."""
    import torch
    from helpers import SIM_PARAMS, v2p_dual_cfg
    from vid2player3d_b200.tasks import PhysicsMVAEControllerDual
    torch.manual_seed(10)
    cfg = v2p_dual_cfg(envs)
    cfg["env"]["motion_player"] = "stream"
    env = PhysicsMVAEControllerDual(cfg, SIM_PARAMS, 1, "cuda", device_index, True)
    dev = env.device
    env.reset()
    acts = [torch.clamp(torch.randn(envs, env.num_actions, device=dev), -5, 5) for _ in range(8)]
    stats = {"resets": 0}

    def run(n):
        for i in range(n):
            env.step(acts[i % 8])
            done = env.reset_buf.nonzero(as_tuple=False).flatten()
            stats["resets"] += len(done)
            env.reset(done)
    run(warmup)
    env.enable_cuda_graph()
    run(warmup)
    torch.cuda.synchronize()
    stats["resets"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    task = env._physics_player.task
    a75 = torch.zeros(envs, task.num_actions, device=dev)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(20):
        for h in task._envs:
            h.step(a75)
    p1.record()
    torch.cuda.synchronize()
    return {"env_steps_per_s": envs * steps / (ms * 1e-3), "ms_per_step": ms / steps, "physics_ms_both_assets": p0.elapsed_time(p1) / 20,
            "steps": steps, "pair_resets_per_step": stats["resets"] / 2 / steps,
            "mode": "one CUDA graph per high-level step (2 physics launches) + reset(done ids) = nonzero() host sync, mask fill, one reset-graph replay",
            "workload": f"vid2player federer_djokovic dual: {envs} paired envs ({envs // 2} rallies), substeps 6, return_w_estimate, "
                        "use_random_ball_target, fix_head_orientation, synthetic incoming-ball table, resident kinematic target stream"}


def ball_tables_workload(reps=3):
    """the reference's offline ball data products at their full sizes, one launch each (tools/perf_ballgen.py has the
    stand-alone version).  out tables: 8 250 000 rows x (60 + 30 x 2) f32; in table: 1 125 000 rows x 50 x 2 f32."""
    import torch
    from vid2player3d_b200 import ball_gen as G
    dev = "cuda:%d" % torch.cuda.current_device()

    def best_ms(fn):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return min(ms)
    vh, vv, vs = G._mesh(G.traj_out_params.VEL_X_RANGE, G.traj_out_params.VEL_Y_RANGE, G.traj_out_params.VSPIN_RANGE, device=dev)
    n = int(vh.shape[0])
    ms_out = best_ms(lambda: G.simulate_without_bounce(vh, vv, vs, device=dev))
    del vh, vv, vs
    hh, vx, vz, sp = G._mesh(G.traj_in_params.HEIGHT_RANGE, G.traj_in_params.VEL_X_RANGE, G.traj_in_params.VEL_Y_RANGE,
                             G.traj_in_params.VSPIN_RANGE, device=dev)
    m = int(hh.shape[0])
    pos = torch.stack([torch.zeros_like(hh), torch.zeros_like(hh), hh], 1)
    vel = torch.stack([torch.zeros_like(hh), vx, vz], 1)
    ms_in = best_ms(lambda: G.simulate(pos, vel, sp, num_frames=50, first_comp=1, device=dev))
    out_bytes = n * (120 * 4 + 12)
    return {"out_tables": {"rows": n, "ms": ms_out, "rows_per_s": n / (ms_out * 1e-3), "algorithmic_bytes": out_bytes,
                           "GBps": out_bytes / 1e9 / (ms_out * 1e-3), "ball_sim_steps_per_s": n * 122 / (ms_out * 1e-3)},
            "in_table": {"rows": m, "ms": ms_in, "rows_per_s": m / (ms_in * 1e-3), "ball_substeps_per_s": m * 50 * 12 / (ms_in * 1e-3)},
            "workload": "offline ball data generators (tennis_ball_out_estimator.py:208-258, tennis_ball_in_estimator.py:82-140) at the "
                        "reference's grid sizes; the reference steps 10 000 balls per Isaac Gym batch from Python"}


def cpu_reference_arm(model, flat, sample_envs, steps, warmup, seed=7):
    """The CPU restatement of the same env step (oracle/physics_ref.c with OpenMP over envs + oracle/ref_port.py
    numpy obs/reward/reset/MoCap), timed on the host cores on a bounded sample of the workload's envs.
    Stand-in for the reference's Isaac Gym CPU pipeline, which cannot run here."""
    import numpy as np
    from helpers import lib_dict
    from oracle import physics_ref, ref_port as R
    from vid2player3d_b200 import abi
    rng = np.random.default_rng(seed)
    n = sample_envs
    ms, verts = abi.pack_model(model, float(model["mass"].sum()) / 90.0)
    cfg = abi.make_cfg(model)
    ml = lib_dict(flat, model)
    mids = rng.integers(0, flat.num_motions(), n)
    t0 = (rng.random(n) * np.maximum(ml["motion_lengths"][mids] - 32 / 30.0, 0)).astype(np.float32)
    st = R.get_motion_state(ml, mids, t0)
    root = np.concatenate([st[0], st[1], st[3], st[4]], -1).astype(np.float64)
    q, qd = st[2].astype(np.float64), st[5].astype(np.float64)
    rbs = np.zeros((n, 24, 13), np.float32)
    rbs[..., 0:3], rbs[..., 3:7] = st[7], st[8]
    orc = R.ImTaskOracle(ml, mids, t0, np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64),
                         np.float32(2) * np.float32(1 / 60), 300, np.full(24, -0.5, np.float32), np.array([3, 7]),
                         np.ones(24, np.float32), ml["motion_bodies"][mids])

    def one_step():
        nonlocal rbs
        a = (rng.random((n, 75)) * 2 - 1).astype(np.float32)
        _, pd, f, t = orc.pre_physics(a, q.astype(np.float32), rbs[:, 0, 3:7])
        rb, _ = physics_ref.control_step(ms, verts, cfg, root, q, qd, pd.astype(np.float64), np.concatenate([f, t], -1).astype(np.float64))
        rbs = rb.astype(np.float32)
        dofs = np.stack([q, qd], -1).astype(np.float32)
        orc.post_physics(rbs, dofs)

    # thread count: "all the host threads it can use" = the count that is fastest for this sample (SMT oversubscription hurts: on the
    # 64-core / 128-thread B200 host 1024 envs per step run at 14 k env-steps/s on 128 threads, 37 k on 64, 49 k on 32)
    one_step()
    best, used = None, physics_ref.set_threads(0)
    for th in sorted({max(1, os.cpu_count() // d) for d in (1, 2, 4, 8)}, reverse=True):
        physics_ref.set_threads(th)
        one_step()
        t0 = time.perf_counter()
        one_step()
        one_step()
        el = time.perf_counter() - t0
        if best is None or el < best:
            best, used = el, th
    physics_ref.set_threads(used)
    cpu_reference_arm.threads = used
    for _ in range(warmup):
        one_step()
    t_start = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t_start
    return n * steps / dt, dt / steps * 1e3


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON: anything a library prints there meanwhile (NCCL prints its version banner on
    # stdout at init) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    workload = (f"embodied_pose amass_im: {args.envs} envs/GPU, SMPL humanoid 24 bodies/69 dof, synthetic MoCap 64x300 frames, "
                f"random policy, reset(all) every {HORIZON} steps")

    if os.environ.get("OMP_NUM_THREADS") == "1" and "TORCHELASTIC_RUN_ID" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(cores)   # torchrun's default of 1 would cripple the CPU restatement (OpenMP over envs)
    if args.impl == "reference":
        if rank != 0:
            return
        from vid2player3d_b200 import model_compiler, motion_lib
        from oracle import physics_ref
        physics_ref.build()
        model = model_compiler.load_compiled("smpl_mesh_humanoid_amass_v1")
        flat = motion_lib.synthetic(model, num_motions=64, num_frames=300, seed=7, sigma=0.05)
        steps, warm = min(args.steps, 40), min(args.warmup, 3)
        v, ms_step = cpu_reference_arm(model, flat, args.cpu_sample_envs, steps, warm)
        sample = f"{args.cpu_sample_envs} of {args.envs} envs x {steps} steps, OpenMP over envs + numpy"
        emit({
            "impl": "reference", "metric": "env-steps/sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 physics / f32 task logic", "data": "synthetic",
            "config": {"workload": workload, "note": "CPU restatement of the same step (oracle/); the reference's Isaac Gym CPU "
                       "pipeline cannot be installed here (closed binary, py3.8) - stand-in, labelled as such"},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": getattr(cpu_reference_arm, "threads", cores), "kind": "port",
                             "sample": sample, "host_logical_cpus": cores},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    model, flat, task, vec = build_workload(args.envs, local_rank, 7 + rank)  # seed += rank like run.py:37
    dev = task.device
    N, K, W = args.envs, args.steps, args.warmup
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    pool = [torch.rand(N, task.num_actions, device=dev, generator=gen) * 2 - 1 for _ in range(16)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (`value`) ----------------
    def run(nsteps, timed):
        evs = []
        for i in range(nsteps):
            if timed:
                flush.zero_()  # L2 flush between timed iterations (outside the event pair)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if i % HORIZON == 0:
                task.reset()
            task.step(pool[i % len(pool)])
            if timed:
                e1.record()
                evs.append((e0, e1, i % HORIZON == 0))
        return evs

    task.reset()
    run(max(W, 3), False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = task._env.launch_count
    try:
        task._env.set_kernel_timing(True)       # CUDA event pair around the physics launch of every step, on the launching stream
    except Exception:
        pass
    t_wall = time.perf_counter()
    evs = run(K, True)
    barrier()
    wall = time.perf_counter() - t_wall
    launches = task._env.launch_count - launches0
    try:
        phys_ms, phys_n = task._env.kernel_ms()
        task._env.set_kernel_timing(False)
    except Exception:
        phys_ms, phys_n = 0.0, 0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b, _ in evs]
    total_ms = sum(step_ms)
    plain = [m for m, (_, _, r) in zip(step_ms, evs) if not r]
    kernel_ms = sum(plain) / max(len(plain), 1)  # steps without a reset = exactly one step_kernel launch

    # back-to-back (hot L2), single event pair: informational
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(K, False)
    e1.record()
    barrier()
    hot_ms = e0.elapsed_time(e1)

    # ---------------- end-to-end arm through the VecTask API with host buffers ----------------
    host_actions = [p.cpu().pin_memory() for p in pool[:4]]
    host_rew = torch.empty(N, dtype=torch.float32).pin_memory()
    host_reset = torch.empty(N, dtype=torch.long).pin_memory()
    dev_act = torch.empty(N, task.num_actions, device=dev)

    def run_e2e(nsteps):
        for i in range(nsteps):
            if i % HORIZON == 0:
                vec.reset()
            dev_act.copy_(host_actions[i % len(host_actions)], non_blocking=True)
            obs, rew, reset, _ = vec.step(dev_act)
            host_rew.copy_(rew, non_blocking=True)
            host_reset.copy_(reset, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the caller consumes rew/reset on the host every step

    run_e2e(max(W // 4, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(K)
    e1.record()
    barrier()
    e2e_eager_ms = e0.elapsed_time(e1)
    # same loop with the wrapper's step captured as one CUDA graph (VecTaskPython.enable_cuda_graph): the host pays one graph
    # launch per step instead of ~8 launches through PyTorch / ctypes; copies, reset and the per-step host sync are unchanged
    vec.enable_cuda_graph(dev_act)
    run_e2e(max(W // 4, 3))
    barrier()
    e0.record()
    run_e2e(K)
    e1.record()
    barrier()
    e2e_ms = min(e2e_eager_ms, e0.elapsed_time(e1))

    # max over ranks
    t = torch.tensor([total_ms, hot_ms, e2e_ms, kernel_ms, e2e_eager_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, hot_ms, e2e_ms, kernel_ms, e2e_eager_ms = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_envs = N * world
    value = total_envs * K / (total_ms * 1e-3)
    peak, peak_src = peaks()
    achieved = ALGO_BYTES_PER_ENV_STEP * N / (kernel_ms * 1e-3) / 1e9  # per-GPU, dominant kernel = step_kernel
    out = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload, "envs_per_gpu": N, "sim": "dt 1/60 x controlFreqInv 2 x substeps 2", "obs": 461,
                   "l2": "flushed (256 MiB memset) between timed steps; per-step CUDA events summed",
                   "value_hot_l2_back_to_back": total_envs * K / (hot_ms * 1e-3), "wall_s_timed_loop": wall,
                   "target_env_steps_per_s_1gpu": 4.0e6},
        "clocks": clocks,
        "e2e": {"value": total_envs * K / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": N * task.num_actions * 4,
                "d2h_bytes_per_step": N * 4 + N * 8,
                "api": "VecTaskPythonWrapper.step/reset (step captured as a CUDA graph: enable_cuda_graph), pinned host actions in, reward+reset "
                       "out, host sync every step", "value_eager_launches": total_envs * K / (e2e_eager_ms * 1e-3)},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                     "kernel": ("step_kernel" if os.environ.get("B200ENV_KERNEL") == "lane" else
                                "step_kernel_packed<fused>" if os.environ.get("B200ENV_SPLIT") == "0" else
                                "one env step = pre_kernel + step_kernel_packed<split> (dominant, ~75 %) + post_kernel"), "kernel_ms": kernel_ms, "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP,
                     "dominant_kernel": {"name": "step_kernel_packed<split>", "ms": phys_ms, "launches_timed": phys_n,
                                         "algorithmic_bytes_per_env": 3044, "GBps": (3044 * N / (phys_ms * 1e-3) / 1e9) if phys_ms > 0 else None,
                                         "note": "CUDA event pair around this launch inside b200env_step (b200env_set_kernel_timing); its share of "
                                                 "the step's algorithmic bytes: state rows + PD targets + wrench in, state / rigid-body / contact rows out"},
                     "peak_source": peak_src,
                     "fp32": {"flop_per_env_step": FP32_FLOP_PER_ENV_STEP, "achieved_tflops": FP32_FLOP_PER_ENV_STEP * N / (kernel_ms * 1e-3) / 1e12,
                              "note": "from the ncu instruction mix (profiles/r1h_step_kernel_packed_ncu.md); B200 non-tensor FP32 peak ~75 TFLOP/s"},
                     "note": "latency / FP32-issue bound along the 9-level kinematic chain, not HBM bound (DESIGN.md 5)"},
    }
    if not args.no_federer and world == 1:
        try:
            out["config"]["federer"] = federer_workload(N, local_rank)
        except Exception as ex:  # secondary measurement must never break the contract line
            out["config"]["federer"] = {"error": repr(ex)[:200]}
        try:
            out["config"]["dual"] = dual_workload(N, local_rank)
        except Exception as ex:
            out["config"]["dual"] = {"error": repr(ex)[:200]}
        try:
            out["config"]["ball_tables"] = ball_tables_workload()
        except Exception as ex:
            out["config"]["ball_tables"] = {"error": repr(ex)[:200]}
    if not args.no_cpu_baseline and world == 1:   # contract: the CPU baseline is timed at N = 1 only
        from oracle import physics_ref
        physics_ref.build()
        v, _ = cpu_reference_arm(model, flat, args.cpu_sample_envs, 24, 2)
        out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": getattr(cpu_reference_arm, "threads", cores), "kind": "port",
                               "host_logical_cpus": cores,
                               "sample": f"{args.cpu_sample_envs} of {N} envs x 24 steps (OpenMP physics restatement + numpy task logic)"}
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
