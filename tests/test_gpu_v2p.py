"""GPU parity of the vid2player kernels (include/b200env_v2p.h) against the fixtures produced by executing the
reference's own vid2player code (tests/golden/v2p_*.npz)."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a, dtype=None):
    t = torch.tensor(a, device=DEV)
    return t.to(dtype).contiguous() if dtype is not None else t.contiguous()


def close(a, b, tol=1e-5, msg=""):
    np.testing.assert_allclose(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a, b, rtol=0, atol=tol, err_msg=msg)


def test_smpl_to_sim_golden():
    from vid2player3d_b200 import native_v2p as V
    g = golden("v2p_smpl_to_sim.npz")
    n = g["root0"].shape[0]
    mk = lambda: dict(root_rot=torch.zeros(n, 4, device=DEV), dof_pos=torch.zeros(n, 69, device=DEV), root_vel=torch.zeros(n, 3, device=DEV),  # noqa: E731
                      root_ang_vel=torch.zeros(n, 3, device=DEV), dof_vel=torch.zeros(n, 69, device=DEV),
                      rb_pos=torch.zeros(n, 24, 3, device=DEV), rb_rot=torch.zeros(n, 24, 4, device=DEV))
    rest, par, s2m = T(g["rest"]), T(g["parents"], torch.int32), T(g["smpl_2_mujoco"], torch.int32)
    a = mk()
    V.smpl_to_sim(T(g["root0"]), T(g["rotmat0"]), rest, par, s2m, float(g["dt"]), a)
    for k in a:
        close(a[k], g["a_" + k], 2e-5, k)
    b = mk()
    V.smpl_to_sim(T(g["root1"]), T(g["rotmat1"]), rest, par, s2m, float(g["dt"]), b, prev_root_pos=T(g["a_root_pos"]), prev_rb_rot=T(g["a_rb_rot"]))
    for k in b:
        tol = {"dof_vel": 2e-3, "root_ang_vel": 6e-2}.get(k, 2e-5)   # finite differences: /dt and /dt^2 amplify 1e-7 rounding
        close(b[k], g["b_" + k], tol, k)


def test_ball_aero_and_reset_golden():
    from vid2player3d_b200 import native_v2p as V
    g = golden("v2p_ball.npz")
    for s in (2, 6):
        bs = T(g[f"s{s}_ball_states"])
        n = bs.shape[0]
        hb, now = T(g[f"s{s}_has_bounce_in"]), torch.zeros(n, dtype=torch.bool, device=DEV)
        bpos, force = torch.zeros(n, 3, device=DEV), torch.zeros(n, 3, device=DEV)
        V.ball_aero(bs, hb, now, bpos, force, s, 5.0)
        close(force, g[f"s{s}_force"], 1e-6)
        assert np.array_equal(hb.cpu().numpy(), g[f"s{s}_has_bounce"]) and np.array_equal(now.cpu().numpy(), g[f"s{s}_has_bounce_now"])
        close(bpos, g[f"s{s}_bounce_pos"], 0)
    N = 64
    ids, pidx = T(g["reset_ids"]), T(g["reset_pool_index"])
    bs = torch.zeros(N, 13, device=DEV)
    bpos, bvel, bounce = torch.zeros(N, 3, device=DEV), torch.zeros(N, 3, device=DEV), torch.ones(N, 3, device=DEV)
    hb, hc = torch.ones(N, dtype=torch.bool, device=DEV), torch.ones(N, dtype=torch.bool, device=DEV)
    traj = torch.zeros(N, 100, 3, device=DEV)
    V.ball_reset(ids, pidx, T(g["pool"]), bs, bpos, bvel, hb, bounce, hc, traj)
    V.ball_reset(ids[:0], pidx[:0], T(g["pool"]), bs, bpos, bvel, hb, bounce, hc, traj)  # empty id list: no-op
    close(bs, g["reset_ball_states"], 2e-5)
    close(traj[ids], g["reset_traj"], 0)
    assert np.array_equal(hb.cpu().numpy(), g["reset_has_bounce"]) and np.array_equal(hc.cpu().numpy(), g["reset_contact"])
    close(bounce, g["reset_bounce_pos"], 0)
    close(bpos[ids], g["reset_ball_states"][g["reset_ids"], 0:3], 0)


def test_update_state_golden():
    from vid2player3d_b200 import native_v2p as V
    g = golden("v2p_update_state.npz")
    for grip in ("eastern", "semi_western"):
        rbs, root, ball = T(g[f"{grip}_rbs"]), T(g[f"{grip}_root_states"]), T(g[f"{grip}_ball_states"])
        n = rbs.shape[0]
        t = dict(has_contact=T(g[f"{grip}_contact_in"]), has_contact_now=torch.zeros(n, dtype=torch.bool, device=DEV),
                 ball_vel=T(g[f"{grip}_prev_ball_vel"]), ball_vspin=torch.zeros(n, device=DEV))
        for k in ("root_pos", "root_vel", "racket_pos", "racket_vel", "racket_normal", "ball_pos"):
            t[k] = torch.zeros(n, 3, device=DEV)
        V.update_state(n, 26, rbs, root, 13, ball, 13, t, grip=grip)
        for k in ("root_pos", "root_vel", "racket_pos", "racket_vel", "racket_normal", "ball_pos", "ball_vel", "ball_vspin"):
            close(t[k], g[f"{grip}_{k}"], 2e-6, k)
        assert np.array_equal(t["has_contact"].cpu().numpy(), g[f"{grip}_contact"])
        assert np.array_equal(t["has_contact_now"].cpu().numpy(), g[f"{grip}_contact_now"])


@pytest.mark.parametrize("rtype", ["reach", "return", "return_w_estimate"])
def test_controller_post_golden(rtype):
    from vid2player3d_b200 import native_v2p as V
    g = golden("v2p_controller.npz")
    N = g["rbs"].shape[0]
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device=DEV, dtype=dt)  # noqa: E731
    obs = z(N, 257)
    t = dict(rigid_body_state=T(g["rbs"]), ball_states=T(g["ball_states"]), root_pos=T(g["p_root_pos"]), root_vel=T(g["p_root_vel"]),
             racket_pos=T(g["p_racket_pos"]), racket_normal=T(g["p_racket_normal"]), ball_pos=T(g["p_ball_pos"]),
             has_contact=T(g["p_has_racket_ball_contact"]), has_contact_now=T(g["p_has_racket_ball_contact_now"]),
             has_bounce=T(g["p_has_bounce"]), has_bounce_now=T(g["p_has_bounce_now"]), bounce_pos=T(g["p_bounce_pos"]),
             ball_traj=T(g["ball_traj"]), target_bounce_pos=T(g["target_bounce_pos"]), phase=T(g["phase"]), swing_type=T(g["swing_type"]),
             swing_type_cycle=T(g["swing_type_cycle"]), tar_action=T(g["tar_action"]), tar_time=T(g["tar_time"]),
             tar_time_total=T(g["tar_time_total"]), progress_buf=T(g["progress"]), est_x=T(g["est_x"]), est_y=T(g["est_y"]),
             bounce_in=z(N, dt=torch.bool), est_bounce_in=z(N, dt=torch.bool), reset_reaction=z(N, dt=torch.bool),
             reset_recovery=z(N, dt=torch.bool), est_bounce_pos=z(N, 3), est_bounce_time=z(N), est_max_height=z(N), distance=z(N),
             obs_buf=obs, rew_buf=z(N), sub_rewards=z(N, 2), reset_buf=z(N, dt=torch.long), terminate_buf=z(N, dt=torch.long))
    cfg = dict(n=N, bodies_per_env=25, ball_stride=13, racket_body=24, num_obs=257, obs_traj_len=10, use_target=1,
               reward_type=V.REWARD_TYPES[rtype], early_termination=1, max_episode_length=300, est_nx=60, est_ny=30, scale_pos=5.0,
               scale_phase=10.0, scale_bounce_pos=0.05, scale_bounce_time=0.1, w_pos=0.5, w_ball_pos=0.5,
               court_min=g["court_min"], court_max=g["court_max"], est_params=g["est_params"].reshape(-1))
    V.controller_post(cfg, t)
    torch.cuda.synchronize()
    close(obs, g["obs"], 2e-6)
    assert np.array_equal(t["bounce_in"].cpu().numpy(), g["bounce_in"])
    close(t["est_bounce_pos"], g["est_bounce_pos"], 1e-5); close(t["est_bounce_time"], g["est_bounce_time"], 1e-6)
    close(t["est_max_height"], g["est_max_height"], 1e-5)
    assert np.array_equal(t["est_bounce_in"].cpu().numpy(), g["est_bounce_in"])
    close(t["rew_buf"], g[f"rew_{rtype}"], 1e-5)
    ns = g[f"sub_{rtype}"].shape[1]
    close(t["sub_rewards"][:, :ns], g[f"sub_{rtype}"], 1e-5)
    close(t["distance"], g["distance"], 1e-6)
    if rtype == "return_w_estimate":   # the golden _compute_reset ran with this reward type and a NaN row in obs
        t["root_pos"][5, 0] = float("nan")   # reproduces obs row 5 carrying a NaN (root_pos feeds obs[0])
        V.controller_post(cfg, t)
        torch.cuda.synchronize()
        for k, gk in (("reset_buf", "reset"), ("terminate_buf", "terminate"), ("reset_reaction", "reset_reaction"), ("reset_recovery", "reset_recovery")):
            got = t[k].cpu().numpy()
            want = g[gk]
            assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), k


def test_controller_end_to_end():
    """config-3 style rollout (synthetic motion generator, zero-residual low-level policy): 150 high-level steps"""
    from helpers import SIM_PARAMS, v2p_cfg
    from oracle import ref_port_v2p as V
    from vid2player3d_b200.tasks import PhysicsMVAEController
    torch.manual_seed(0)
    N = 256
    env = PhysicsMVAEController(v2p_cfg(N), SIM_PARAMS, 1, "cuda", 0, True)
    assert env.num_obs == 257 and env.num_actions == 35
    env.reset()
    task = env._physics_player.task
    torch.cuda.synchronize()
    assert torch.isfinite(env.obs_buf).all()
    y0 = task._ball_pos[:, 1].clone()
    assert (y0 > 11).all() and (task._ball_vel[:, 1] < -15).all()          # launched from the far side towards the player
    bounced, resets, hits = 0, 0, 0
    for step in range(150):
        a = torch.clamp(torch.randn(N, 35, device=DEV), -5, 5)
        env.step(a)
        done = env.reset_buf.nonzero(as_tuple=False).flatten()
        resets += len(done)
        env.reset(done)                                                      # agent loop: env_reset(done_indices)
        bounced = max(bounced, int(task._has_bounce.sum()))
        hits += int(task._has_racket_ball_contact_now.sum())
    torch.cuda.synchronize()
    assert torch.isfinite(env.obs_buf).all() and torch.isfinite(env.rew_buf).all()
    assert bounced > N // 4                       # balls reached the ground on the player's side
    assert (env.progress_buf > 0).any() and (env._num_reset_reaction > 1).any()   # reaction tasks were re-armed (tar_time FSM)
    close(env.obs_buf[:, 0:3], task._root_pos.cpu().numpy(), 0)
    # the fused post kernel agrees with the numpy oracle on the live GPU state
    g = lambda t: t.detach().cpu().numpy()  # noqa: E731
    rbs = g(task._rigid_body_state.view(N, 26, 13))
    obs = V.controller_obs(rbs[:, :25], g(task._root_pos), g(task._root_vel), g(task._racket_normal), g(env._ball_traj),
                           g(env._target_bounce_pos), 10)
    close(env.obs_buf, obs, 2e-6)
    qn = task._rigid_body_rot.norm(dim=-1)
    assert (qn - 1).abs().max() < 1e-4
    assert (task._ball_pos[:, 2] >= 0.0319).all()  # never below the ground


def test_racket_hit_is_detected():
    """aim the ball at the racket face: the swept impact fires, the velocity-jump detector (substeps > 2) or the exact flag sees it"""
    from helpers import SIM_PARAMS, v2p_cfg
    from vid2player3d_b200.tasks import PhysicsMVAEController
    torch.manual_seed(1)
    N = 64
    env = PhysicsMVAEController(v2p_cfg(N), SIM_PARAMS, 1, "cuda", 0, True)
    env.reset()
    task = env._physics_player.task
    env.step(torch.zeros(N, 35, device=DEV))      # one step so the racket row comes from the simulated FK
    n = task._racket_normal
    centre = task._racket_pos + 0.02125 * n
    ball = task._ball_root_states
    ball[:, 0:3] = centre + 0.25 * n
    ball[:, 7:10] = task._racket_vel - 25.0 * n
    ball[:, 10:13] = 0
    task._ball_vel.copy_(ball[:, 7:10])
    v_before = ball[:, 7:10].clone()
    env.step(torch.zeros(N, 35, device=DEV))
    torch.cuda.synchronize()
    assert task._racket_hit_now.float().mean() > 0.9
    dv = (task._ball_root_states[:, 7:10] - v_before).norm(dim=-1)
    assert (dv[task._racket_hit_now] > 20).all()   # restitution 0.9: the normal velocity is reversed
