"""Timing of the offline ball data generators (include/b200ball.h) at the reference's full sizes: python tools/perf_ballgen.py [reps]
   out tables: 550 x 150 x 100 = 8 250 000 rows -> [rows,60] + [rows,30,2] f32 (3.96 GB)      tennis_ball_out_estimator.py:208-258
   in table:   15 x 50 x 30 x 50 = 1 125 000 rows -> [rows,50,2] f32 (0.45 GB)                tennis_ball_in_estimator.py:82-140
   pool:       10 000 launches x 100 frames per reset() of the generator                       tennis_ball.py:278-311
CUDA events around the kernel launch only (launch arrays and zeroed outputs are prepared before)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from vid2player3d_b200 import ball_gen as G


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return min(ms), sum(ms) / len(ms)


def main(reps=5):
    dev = "cuda:0"
    out = {}
    vh, vv, vs = G._mesh(G.traj_out_params.VEL_X_RANGE, G.traj_out_params.VEL_Y_RANGE, G.traj_out_params.VSPIN_RANGE, device=dev)
    n = vh.shape[0]
    best, avg = timed(lambda: G.simulate_without_bounce(vh, vv, vs), reps)        # includes the zero-fill of the 3.96 GB of outputs
    bytes_out = n * (60 + 60) * 4 + n * 12
    out["out_tables"] = dict(rows=n, ms_best=best, ms_avg=avg, rows_per_s=n / (best * 1e-3), algorithmic_GB=bytes_out / 1e9,
                             GBps=bytes_out / 1e9 / (best * 1e-3), sim_steps_per_s=n * 122 / (best * 1e-3))
    del vh, vv, vs
    hh, vx, vz, sp = G._mesh(G.traj_in_params.HEIGHT_RANGE, G.traj_in_params.VEL_X_RANGE, G.traj_in_params.VEL_Y_RANGE,
                             G.traj_in_params.VSPIN_RANGE, device=dev)
    m = hh.shape[0]
    pos = torch.stack([torch.zeros_like(hh), torch.zeros_like(hh), hh], 1)
    vel = torch.stack([torch.zeros_like(hh), vx, vz], 1)
    best, avg = timed(lambda: G.simulate(pos, vel, sp, num_frames=50, first_comp=1), reps)
    b = m * (50 * 2 * 4 + 28 + 21)
    out["in_table"] = dict(rows=m, ms_best=best, ms_avg=avg, rows_per_s=m / (best * 1e-3), GBps=b / 1e9 / (best * 1e-3),
                           substeps_per_s=m * 50 * 12 / (best * 1e-3))
    gen = G.TennisBallGeneratorB200({}, need_reset=False, generator=torch.Generator().manual_seed(0))
    best, avg = timed(gen.reset, reps)
    out["pool_reset_10000"] = dict(ms_best=best, ms_avg=avg, kept=len(gen.traj_pool))
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
