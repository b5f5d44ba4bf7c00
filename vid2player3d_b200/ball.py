"""Tennis-ball data products the vid2player envs consume: the incoming-trajectory pool and the outgoing-ball
estimator tables.  The reference builds them offline with Isaac Gym
(vid2player/utils/tennis_ball.py:221-394, tennis_ball_out_estimator.py:21-121,208-258) and does not ship them
(data/ball_traj/*.npy are git-ignored), so this module fabricates physically consistent ones with the same
ball model the CUDA step integrates (drag + Magnus lift + restitution bounce), in the reference's file formats:

  pool    [P, 307]  = launch_pos 3 | launch_vel 3 | vspin 1 | trajectory 100 x 3 @ 30 Hz   (tennis_ball.py:424-431)
  out_x   [rows, NX] height above launch as a function of horizontal distance (TRAJ_X grid)
  out_y   [rows, NY, 2] (horizontal distance to the bounce, time to the bounce) per launch height (TRAJ_Y grid)
  rows indexed by (vel_x, vel_y, vspin) exactly like TennisBallOutEstimator.get_ball_traj_out_index (:126-144).
"""
import math

import numpy as np

R, M, RHO, G = 0.032, 0.057, 1.21, 9.81
KF = RHO * math.pi * R * R / 2
CD = 0.55


def _aero(v, w, spin_scale):
    vs = np.linalg.norm(v, axis=-1, keepdims=True)
    vs = np.where(vs == 0, 1.0, vs)
    vn = v / vs
    vt = np.cross(vn, np.array([0.0, 0.0, -1.0]))
    vspin = np.linalg.norm(w, axis=-1, keepdims=True) / (2 * math.pi)
    cl = 1.0 / (2.0 + np.abs(vs / (vspin * spin_scale + 1e-6)))
    cl = np.where(vspin > 0, -cl, cl)
    return -KF * CD * vs * v - KF * cl * vs ** 2 * np.cross(vt, vn)


def integrate(pos, vel, vspin, n_frames=100, fps=30, substeps=12, spin_scale=5.0, e=0.7, mu=0.6, bounce=True):
    """Trajectories [n, n_frames, 3] sampled at `fps`; angular velocity = vspin*2pi*normalize(v x -z) (topspin convention of
    _reset_balls, humanoid_smpl_im_mvae.py:508-509)."""
    pos, vel = np.array(pos, np.float64), np.array(vel, np.float64)
    c = np.cross(vel, np.array([0.0, 0.0, -1.0]))
    w = np.asarray(vspin)[:, None] * 2 * math.pi * c / np.maximum(np.linalg.norm(c, axis=-1, keepdims=True), 1e-12)
    h = 1.0 / fps / substeps
    out = np.zeros((len(pos), n_frames, 3))
    for f in range(n_frames):
        out[:, f] = pos
        for _ in range(substeps):
            vel = vel + h * (_aero(vel, w, spin_scale) / M + np.array([0.0, 0.0, -G]))
            pos = pos + h * vel
            if bounce:
                hit = (pos[:, 2] < R) & (vel[:, 2] < 0)
                if hit.any():
                    vel[hit, 2] *= -e
                    vel[hit, :2] *= (1 - 0.4 * mu)
                    pos[hit, 2] = R
    return out


def synthetic_pool(P=10000, seed=10, spin_scale=5.0):
    """launch distribution of the reference's generator (tennis_ball.py:278-284): origin x in [-4,4], y in [12,13],
    z in [1,1.5]; speed 28-30 m/s towards -y; elevation 5-15 deg; spin 5-10 rps."""
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.uniform(-4, 4, P), rng.uniform(12, 13, P), rng.uniform(1, 1.5, P)], -1)
    speed, elev = rng.uniform(28, 30, P), np.deg2rad(rng.uniform(5, 15, P))
    tx = rng.uniform(-3.5, 3.5, P)                      # aim point on the player's side
    dirxy = np.stack([tx - pos[:, 0], -11.0 - pos[:, 1]], -1)
    dirxy /= np.linalg.norm(dirxy, axis=-1, keepdims=True)
    vel = np.concatenate([dirxy * (speed * np.cos(elev))[:, None], (speed * np.sin(elev))[:, None]], -1)
    spin = rng.uniform(5, 10, P)
    traj = integrate(pos, vel, spin, spin_scale=spin_scale)
    return np.concatenate([pos, vel, spin[:, None], traj.reshape(P, 300)], -1).astype(np.float32)


COARSE_PARAMS = dict(VEL_X=(10.0, 65.0, 1.0), VEL_Y=(-5.0, 10.0, 1.0), VSPIN=(-10.0, 10.0, 2.0), TRAJ_X=(0.0, 30.0, 0.5), TRAJ_Y=(0.0, 3.0, 0.1))


_TABLE_CACHE = {}


def synthetic_out_tables(params=None, spin_scale=5.0):
    """Outgoing-ball tables on a coarse grid (the reference's full grid is 550x150x100 rows ~ 4 GB).  Returns
    (out_x [rows,NX], out_y [rows,NY,2], params_array [5,3])."""
    key = (repr(sorted((params or {}).items())), float(spin_scale))
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    p = dict(COARSE_PARAMS)
    p.update(params or {})
    axes = [np.arange(lo, hi - 1e-9, st) for lo, hi, st in (p["VEL_X"], p["VEL_Y"], p["VSPIN"])]
    vx, vy, vs = [a.ravel() for a in np.meshgrid(*axes, indexing="ij")]
    rows = len(vx)
    xs = np.arange(p["TRAJ_X"][0], p["TRAJ_X"][1] - 1e-9, p["TRAJ_X"][2])
    ys = np.arange(p["TRAJ_Y"][0], p["TRAJ_Y"][1] - 1e-9, p["TRAJ_Y"][2])
    # 2-D flight (horizontal distance d, height z) without bounce; spin sign = top(+)/back(-) spin
    h, T = 1.0 / 360.0, 3.0
    n = int(T / h)
    d, z = np.zeros(rows), np.zeros(rows)
    vd, vz = vx.copy(), vy.copy()
    D, Z = np.zeros((rows, n)), np.zeros((rows, n))
    for i in range(n):
        D[:, i], Z[:, i] = d, z
        sp = np.hypot(vd, vz)
        cl = 1.0 / (2.0 + np.abs(sp / (np.abs(vs) * spin_scale + 1e-6)))
        lift = -np.sign(vs) * KF * cl * sp ** 2 / M          # topspin pushes down
        ad = -KF * CD * sp * vd / M - lift * vz / np.maximum(sp, 1e-9)
        az = -G - KF * CD * sp * vz / M + lift * vd / np.maximum(sp, 1e-9)
        vd, vz = vd + h * ad, vz + h * az
        d, z = d + h * vd, z + h * vz
    out_x = np.stack([Z[np.arange(rows), np.minimum((np.abs(D - x) .argmin(axis=1)), n - 1)] for x in xs], -1)
    out_y = np.zeros((rows, len(ys), 2))
    t = np.arange(n) * h
    for j, y0 in enumerate(ys):
        below = (Z + y0 - R) < 0
        idx = np.where(below.any(axis=1), below.argmax(axis=1), n - 1)
        out_y[:, j, 0], out_y[:, j, 1] = D[np.arange(rows), idx], t[idx]
    arr = np.array([p["VEL_X"], p["VEL_Y"], p["VSPIN"], p["TRAJ_X"], p["TRAJ_Y"]], np.float64)
    _TABLE_CACHE[key] = (out_x.astype(np.float32), out_y.astype(np.float32), arr)
    return _TABLE_CACHE[key]


IN_PARAMS = dict(HEIGHT=(0.5, 2.0, 0.1), VEL_X=(25.0, 30.0, 0.1), VEL_Y=(5.0, 8.0, 0.1), VSPIN=(5.0, 10.0, 0.1))   # traj_in_params
IN_PARAMS_COARSE = dict(HEIGHT=(0.5, 2.0, 0.25), VEL_X=(25.0, 30.0, 0.5), VEL_Y=(5.0, 8.0, 0.5), VSPIN=(5.0, 10.0, 1.0))


def synthetic_in_table(params=None, spin_scale=5.0):
    """Incoming-ball table of dual mode in the reference's file format (utils/tennis_ball_in_estimator.py:82-140):
    rows over (height, horizontal speed, vertical speed, spin) in C order, each 50 frames @ 30 Hz of (distance along the hit
    direction, height) for a ball launched straight out from (0, 0, height).  Returns (table [rows,50,2] f32, params [4,3] f64).
    Default = a coarse grid (2 520 rows); pass IN_PARAMS for the shipped 1 125 000-row grid (~450 MB)."""
    p = dict(IN_PARAMS_COARSE)
    p.update(params or {})
    key = ("in", repr(sorted(p.items())), float(spin_scale))
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    axes = [np.arange(*p[k]) for k in ("HEIGHT", "VEL_X", "VEL_Y", "VSPIN")]
    hh, vx, vz, vs = [a.ravel() for a in np.meshgrid(*axes, indexing="ij")]
    rows = len(hh)
    out = np.zeros((rows, 50, 2), np.float32)
    for s in range(0, rows, 200000):
        e = min(rows, s + 200000)
        pos = np.stack([np.zeros(e - s), np.zeros(e - s), hh[s:e]], -1)
        vel = np.stack([np.zeros(e - s), vx[s:e], vz[s:e]], -1)
        out[s:e] = integrate(pos, vel, vs[s:e], n_frames=50, spin_scale=spin_scale)[:, :, 1:]
    arr = np.array([p["HEIGHT"], p["VEL_X"], p["VEL_Y"], p["VSPIN"]], np.float64)
    _TABLE_CACHE[key] = (out, arr)
    return _TABLE_CACHE[key]
