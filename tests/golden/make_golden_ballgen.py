"""Golden fixture for the offline ball data generators, produced by EXECUTING the reference's own
`simulate` (vid2player/utils/tennis_ball.py:113-218) and `simulate_without_bounce`
(vid2player/utils/tennis_ball_out_estimator.py:21-121) on CPU.  Isaac Gym is absent, so the `gym` object handed to
them is a fake whose `simulate()` advances the root-state tensor with OUR ball model
(oracle/ref_port_ballgen.py::BallWorld); everything the reference does around that call - force evaluation, flags,
sampling, grid resampling - runs unmodified.  Run in the build container only:

    python tests/golden/make_golden_ballgen.py      ->  tests/golden/v2p_ballgen.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refenv  # noqa: E402

_refenv.setup("vid2player")

from utils import tennis_ball as TB  # noqa: E402
from utils import tennis_ball_out_estimator as TE  # noqa: E402

from oracle import ref_port_ballgen as P  # noqa: E402

torch.set_num_threads(1)


class FakeGym:
    """the 7 gym calls `simulate` / `simulate_without_bounce` make; state = float32 torch tensor [num_env, 13]"""

    def __init__(self, num_env, substeps, ground):
        self.state = torch.zeros(num_env, 13, dtype=torch.float32)
        self.state[:, 2] = 1.0
        self.state[:, 6] = 1.0
        self.force = np.zeros((num_env, 3))
        self.world = P.BallWorld(num_env, substeps=substeps, ground=ground)

    def acquire_actor_root_state_tensor(self, sim):
        return self.state

    def set_actor_root_state_tensor(self, sim, t):
        assert t is self.state or t.data_ptr() == self.state.data_ptr()

    def apply_rigid_body_force_tensors(self, sim, forces, torques, space):
        self.force = forces.detach().numpy().astype(np.float64).copy()

    def simulate(self, sim):
        s = self.state.numpy().astype(np.float64)
        self.world.sim_step(s, self.force)
        self.state.copy_(torch.from_numpy(s.astype(np.float32)))
        self.force[:] = 0   # Isaac Gym clears applied forces after every simulate()

    def fetch_results(self, sim, wait):
        pass

    def refresh_actor_root_state_tensor(self, sim):
        pass


def launches(rng, n):
    """launch distribution of TennisBallGeneratorIsaac.reset (tennis_ball.py:278-296) + a few edge rows"""
    origin = rng.uniform([-4, 12, 1], [4, 13, 1.5], (n, 3))
    bounce = rng.uniform([-3, -10, 0], [3, -7, 0], (n, 3))
    d = bounce[:, :2] - origin[:, :2]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    speed, theta, spin = rng.uniform(28, 30, n), np.deg2rad(rng.uniform(5, 15, n)), rng.uniform(5, 10, n)
    vel = np.stack([speed * np.cos(theta) * d[:, 0], speed * np.cos(theta) * d[:, 1], speed * np.sin(theta)], 1)
    spin[:4] *= -1                         # back-spin launches: exercises the sign flip at the first bounce
    vel[4:8] *= 0.55                       # short balls: bounce before the net -> pass_net False
    vel[8:10, 2] = -6.0                    # driven into the ground early
    return origin.astype(np.float32), vel.astype(np.float32), spin.astype(np.float32)


def main():
    rng = np.random.default_rng(20)
    out = {}
    # ---- simulate(): pool rows (100 frames) and incoming-table rows (50 frames, launched straight out from x = y = 0)
    n, extra = 48, 5
    lp, lv, ls = launches(rng, n)
    gym = FakeGym(n + extra, 6, True)      # more envs than balls: the reference zeroes forces[num_ball:]
    traj, bpos, bidx, pnet = TB.simulate(gym, None, torch.from_numpy(lp), torch.from_numpy(lv), torch.from_numpy(ls),
                                         num_frames=100, substeps=6, spin_scale=5)
    out.update(sim_pos=lp, sim_vel=lv, sim_vspin=ls, sim_traj=traj.numpy(), sim_bounce_pos=bpos.numpy(), sim_bounce_idx=bidx.numpy(),
               sim_pass_net=pnet.numpy())
    h, vx, vz, vs = (g[:: 9973][:40] for g in P.launch_grid_in())
    lp2 = np.stack([np.arange(len(h)) / 1000, np.zeros(len(h)), h], 1).astype(np.float32)
    lv2 = np.stack([np.zeros(len(h)), vx, vz], 1).astype(np.float32)
    gym = FakeGym(len(h), 6, True)
    traj2, _, _, _ = TB.simulate(gym, None, torch.from_numpy(lp2), torch.from_numpy(lv2), torch.from_numpy(vs.copy()), num_frames=50)
    out.update(in_pos=lp2, in_vel=lv2, in_vspin=vs, in_traj=traj2.numpy())
    # substeps == 2 branch (bounce threshold 4R)
    gym = FakeGym(16, 2, True)
    traj3, bpos3, bidx3, pnet3 = TB.simulate(gym, None, torch.from_numpy(lp[:16]), torch.from_numpy(lv[:16]), torch.from_numpy(ls[:16].copy()),
                                             num_frames=100, substeps=2, spin_scale=5)
    out.update(sim2_traj=traj3.numpy(), sim2_bounce_pos=bpos3.numpy(), sim2_bounce_idx=bidx3.numpy(), sim2_pass_net=pnet3.numpy())
    # ---- simulate_without_bounce(): rows of the outgoing-ball tables.  The reference launches at z = 100 (no ground in reach)
    # and then shifts the heights with `traj_all[:, :, 1] -= traj_all[0, 0, 1]` (:83) - an in-place op whose operand is a view
    # of its own output: under torch 2.11 on CPU only element [0, 0] is shifted (the operand reads 0 afterwards) and the
    # table comes out as extrapolation garbage.  What the consumer needs (TennisBallOutEstimator.estimate :164-205) is the
    # height relative to the launch height, so the fixture launches at z = 0 over a world without ground: the shift is then a
    # no-op whatever the aliasing does, and every other line of the function is pinned as it stands.
    a, b, c = P.launch_grid_out()
    pick = np.concatenate([np.arange(0, len(a), 131071), [len(a) - 1, 50, 49, 51, 14950, 8249950]])   # spin 0 / +-0.2, slow and fast rows
    a, b, c = a[pick], b[pick], c[pick]
    m = len(a)
    lp3 = np.zeros((m, 3), np.float32)
    lp3[:, 2] = 0
    lp3[:, 0] = np.arange(m) / 1000
    lv3 = np.stack([np.zeros(m), a, b], 1).astype(np.float32)
    gym = FakeGym(m, 6, False)
    tx, ty = TE.simulate_without_bounce(gym, None, torch.from_numpy(lp3), torch.from_numpy(lv3), torch.from_numpy(c.copy()),
                                        TE.traj_out_params)
    out.update(out_vel=lv3, out_vspin=c, out_x=tx, out_y=ty)
    np.savez_compressed(os.path.join(HERE, "v2p_ballgen.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
