"""ctypes binding of the vid2player entry points (include/b200env_v2p.h) - same library, no CPU fallback."""
import ctypes as C

from . import abi
from .native import _ptr, _stream, lib

SYMBOLS = ["b200v2p_last_error", "b200v2p_smpl_to_sim", "b200v2p_ball_aero", "b200v2p_ball_reset", "b200v2p_ball_in_estimate", "b200v2p_update_state",
           "b200v2p_controller_post", "b200v2p_task_reset", "b200v2p_actor_reset", "b200v2p_fix_head", "b200v2p_pre_step", "b200v2p_stream_gather", "b200v2p_ctrl_reset"]
GRIP_NORMAL = {'eastern': (0.0, 1.0, 0.0), 'semi_western': (0.0, 2.0 ** -0.5, 2.0 ** -0.5)}
REWARD_TYPES = {'reach': 0, 'return': 1, 'return_w_estimate': 2}


def _check(rc):
    if rc != 0:
        L = lib()
        L.b200v2p_last_error.restype = C.c_char_p
        raise RuntimeError(f"b200v2p error {rc}: {L.b200v2p_last_error().decode()}")


def _c(t):
    assert t is None or (t.is_cuda and t.is_contiguous()), "tensors must be contiguous CUDA tensors"
    return _ptr(t)


def smpl_to_sim(root_pos, joint_rotmat, rest, parents, smpl_2_mujoco, dt, out, prev_root_pos=None, prev_rb_rot=None, prev_root_pos_update=None,
                target_root_pos_out=None, only_mask=None):
    """out: dict with root_rot[n,4] dof_pos[n,69] root_vel[n,3] root_ang_vel[n,3] dof_vel[n,69] rb_pos[n,24,3] rb_rot[n,24,4];
    prev_root_pos_update / target_root_pos_out: the root position of this call is also stored there (may alias prev_root_pos)"""
    n = int(root_pos.shape[0])
    num_rest = 1 if rest.dim() == 2 else int(rest.shape[0])      # rest [24,3] or [S,24,3]: env e uses shape e % S
    assert rest.is_contiguous() and rest.shape[-2:] == (24, 3)
    _check(lib().b200v2p_smpl_to_sim(C.c_int32(n), _c(root_pos), _c(joint_rotmat), _c(rest), C.c_int32(num_rest), _c(parents), _c(smpl_2_mujoco), C.c_float(dt),
                                     _c(prev_root_pos), _c(prev_rb_rot), _c(out["root_rot"]), _c(out["dof_pos"]), _c(out["root_vel"]),
                                     _c(out["root_ang_vel"]), _c(out["dof_vel"]), _c(out["rb_pos"]), _c(out["rb_rot"]), _c(prev_root_pos_update),
                                     _c(target_root_pos_out), _c(only_mask), _stream()))


def fix_head(rb_pos, rb_rot, ball_pos, root_pos, joint_rotmat, head_body=13):
    _check(lib().b200v2p_fix_head(C.c_int32(int(rb_pos.shape[0])), _c(rb_pos), _c(rb_rot), C.c_int32(head_body), _c(ball_pos), _c(root_pos),
                                  _c(joint_rotmat), _stream()))


def ball_aero(ball_states, has_bounce, has_bounce_now, bounce_pos, force, substeps, spin_scale, stride=None):
    n = int(has_bounce.shape[0])
    _check(lib().b200v2p_ball_aero(C.c_int32(n), _c(ball_states), C.c_int32(stride or ball_states.stride(0)), _c(has_bounce), _c(has_bounce_now),
                                   _c(bounce_pos), _c(force), C.c_int32(substeps), C.c_float(spin_scale), _stream()))


def ball_reset(env_ids, pool_index, pool, ball_states, ball_pos, ball_vel, has_bounce, bounce_pos, has_contact, traj, stride=None):
    """ball_states: first ball row; `stride` floats between consecutive envs' ball rows (26 when the ball is actor 1 of 2)"""
    _check(lib().b200v2p_ball_reset(C.c_int32(int(env_ids.shape[0])), _c(env_ids), _c(pool_index), _c(pool), _c(ball_states),
                                    C.c_int32(stride or ball_states.stride(0)), _c(ball_pos), _c(ball_vel), _c(has_bounce), _c(bounce_pos),
                                    _c(has_contact), _c(traj), _stream()))


def update_state(n, bodies_per_env, rigid_body_state, root_states, root_stride, ball_states, ball_stride, t, grip='eastern',
                 racket_body=24, wrist_body=22, only_mask=None):
    """grip / racket_body / wrist_body may be 2-sequences (dual_mode 'different': even envs, odd envs)"""
    s = abi.V2PState()
    pair = lambda v: (v[0], v[1], 1) if isinstance(v, (list, tuple)) else (v, v, 0)  # noqa: E731
    g0, g1, d0 = pair(grip)
    r0, r1, d1 = pair(racket_body)
    w0, w1, d2 = pair(wrist_body)
    s.n, s.bodies_per_env, s.ball_stride, s.root_stride, s.racket_body, s.wrist_body = n, bodies_per_env, ball_stride, root_stride, r0, w0
    s.dual, s.racket_body2, s.wrist_body2 = int(bool(d0 or d1 or d2)), r1, w1
    for i in range(3):
        s.grip_normal[i] = GRIP_NORMAL[g0][i]
        s.grip_normal2[i] = GRIP_NORMAL[g1][i]
    s.rigid_body_state, s.root_states, s.ball_states = rigid_body_state.data_ptr(), root_states.data_ptr(), ball_states.data_ptr()
    for k in ("has_contact", "has_contact_now", "root_pos", "root_vel", "racket_pos", "racket_vel", "racket_normal", "ball_pos", "ball_vel",
              "ball_vspin"):
        assert t[k].is_cuda and t[k].is_contiguous(), k
        setattr(s, k, t[k].data_ptr())
    s.only_mask = only_mask.data_ptr() if only_mask is not None else None
    _check(lib().b200v2p_update_state(C.byref(s), _stream()))


def ball_in_estimate(contact_ids, ball_states, stride, table, params, traj, states_in, states_out):
    """TennisBallInEstimator.estimate for the balls of envs `contact_ids`; params [4,3] float64 (HEIGHT, VEL_X, VEL_Y, VSPIN)"""
    import numpy as np
    n = int(contact_ids.shape[0])
    p = np.ascontiguousarray(np.asarray(params, np.float64).reshape(12))
    for x in (contact_ids, table, traj, states_in, states_out):
        assert x.is_cuda and x.is_contiguous()
    assert table.shape[1:] == (50, 2) and traj.shape == (n, 50, 3) and states_in.shape == (n, 13) and states_out.shape == (n, 13)
    _check(lib().b200v2p_ball_in_estimate(C.c_int32(n), _ptr(contact_ids), _ptr(ball_states), C.c_int32(stride), _ptr(table),
                                          C.c_int64(int(table.shape[0])), p.ctypes.data_as(C.c_void_p), _ptr(traj), _ptr(states_in),
                                          _ptr(states_out), _stream()))


def controller_post(cfg, t):
    """cfg: dict of scalars (see abi.V2PCtrl); t: dict of tensors keyed like the struct's pointer fields (est_x/est_y may be None)."""
    c = abi.V2PCtrl()
    for k in ("n", "bodies_per_env", "ball_stride", "racket_body", "num_obs", "obs_traj_len", "use_target", "reward_type",
              "early_termination", "max_episode_length", "est_nx", "est_ny", "scale_pos", "scale_phase", "scale_bounce_pos",
              "scale_bounce_time", "w_pos", "w_ball_pos"):
        setattr(c, k, cfg[k])
    c.obs_only = int(cfg.get("obs_only", 0))
    c.dual = int(cfg.get("dual", 0))
    c.use_history = int(cfg.get("use_history", 0))
    c.advance = int(cfg.get("advance", 0))
    for k in ("court_min", "court_max", "est_params"):
        for i, v in enumerate(cfg[k]):
            getattr(c, k)[i] = float(v)
    scalars = {f for f, _ in abi.V2PCtrl._fields_ if f in cfg} | {"obs_only", "dual", "use_history", "advance"}
    for name, _ in abi.V2PCtrl._fields_:
        if name in scalars:
            continue
        x = t.get(name)
        if x is None:
            assert name in ("est_x", "est_y", "ball_obs", "touch_mask"), name
            setattr(c, name, None)
        else:
            assert x.is_cuda and x.is_contiguous(), name
            setattr(c, name, x.data_ptr())
    _check(lib().b200v2p_controller_post(C.byref(c), _stream()))


def task_reset(cfg, t):
    """cfg: scalars of abi.V2PTaskReset (n, pool_size, ball_stride, bodies_per_env, reaction_nframes, target_mode, target_min/max);
    t: tensors keyed like the struct's pointer fields"""
    r = abi.V2PTaskReset()
    for k in ("n", "pool_size", "ball_stride", "bodies_per_env", "reaction_nframes", "target_mode"):
        setattr(r, k, int(cfg[k]))
    for k in ("target_min", "target_max"):
        for i, v in enumerate(cfg[k]):
            getattr(r, k)[i] = float(v)
    r.obs_traj_len = int(cfg.get("obs_traj_len", 0))
    for name, _ in abi.V2PTaskReset._fields_:
        if name in cfg or name in ("obs_traj_len", "pad_"):
            continue
        x = t.get(name)
        if x is None:
            assert name == "ball_obs", name
            setattr(r, name, None)
            continue
        assert x.is_cuda and x.is_contiguous(), name
        setattr(r, name, x.data_ptr())
    _check(lib().b200v2p_task_reset(C.byref(r), _stream()))


def actor_reset(cfg, t):
    r = abi.V2PActorReset()
    for k in ("n", "num_dof", "bodies_per_env", "root_stride", "racket_body", "racket_parent"):
        setattr(r, k, int(cfg[k]))
    for i, v in enumerate(cfg["racket_offset"]):
        r.racket_offset[i] = float(v)
    if cfg.get("racket_offset2") is not None:
        r.dual = 1
        r.racket_parent2 = int(cfg.get("racket_parent2", cfg["racket_parent"]))
        for i, v in enumerate(cfg["racket_offset2"]):
            r.racket_offset2[i] = float(v)
    for name, _ in abi.V2PActorReset._fields_:
        if name in cfg or name in ("dual", "racket_offset2", "racket_parent2", "pad_"):
            continue
        x = t.get(name)
        if x is None:
            assert name == "mask", name
            setattr(r, name, None)
            continue
        assert x.is_cuda and x.is_contiguous(), name
        setattr(r, name, x.data_ptr())
    _check(lib().b200v2p_actor_reset(C.byref(r), _stream()))


def pre_step(cfg, t):
    """cfg: n, num_actions, num_latent, num_res_dof, random_walk_in_recovery, vae_action_scale, residual_dof_scale, seed;
    t: actions, tar_action, step_counter (int64 [1]), done_counter (int32 [1]), mvae_actions, res_dof_actions (or None)"""
    p = abi.V2PPreStep()
    for k in ("n", "num_actions", "num_latent", "num_res_dof", "random_walk_in_recovery", "seed"):
        setattr(p, k, int(cfg[k]))
    p.vae_action_scale, p.residual_dof_scale = float(cfg["vae_action_scale"]), float(cfg["residual_dof_scale"])
    for k in ("actions", "tar_action", "step_counter", "done_counter", "mvae_actions", "res_dof_actions"):
        x = t.get(k)
        assert x is None or (x.is_cuda and x.is_contiguous()), k
        setattr(p, k, x.data_ptr() if x is not None else None)
    _check(lib().b200v2p_pre_step(C.byref(p), _stream()))


def stream_gather(n, frames, advance, t, reseed_mask=None, seed=0):
    """t: clock (int64 [1]), done_counter (int32 [1]), offset [n] int64, ring_* / live buffers of include/b200env_v2p.h b200v2p_stream_t;
    reseed_mask (bool [n]): the masked envs draw a new offset and re-read their frame, the others are left alone"""
    s = abi.V2PStream()
    s.n, s.frames, s.advance = int(n), int(frames), int(advance)
    s.seed = int(seed)
    s.reseed_mask = reseed_mask.data_ptr() if reseed_mask is not None else None
    for name, _ in abi.V2PStream._fields_[4:]:
        if name in ("reseed_mask", "seed"):
            continue
        x = t[name]
        assert x.is_cuda and x.is_contiguous(), name
        setattr(s, name, x.data_ptr())
    _check(lib().b200v2p_stream_gather(C.byref(s), _stream()))


def ctrl_reset(mask, progress_buf, reset_buf, terminate_buf, num_reset_reaction, distance, num_reset):
    for x in (mask, progress_buf, reset_buf, terminate_buf, num_reset_reaction, distance, num_reset):
        assert x.is_cuda and x.is_contiguous()
    _check(lib().b200v2p_ctrl_reset(C.c_int32(int(mask.shape[0])), _ptr(mask), _ptr(progress_buf), _ptr(reset_buf), _ptr(terminate_buf),
                                    _ptr(num_reset_reaction), _ptr(distance), _ptr(num_reset), _stream()))
