"""ORACLE (test infrastructure, not product code) - numpy restatement of the reference's
pure-PyTorch half of the rollout hot path: observation / reward / reset / MoCap sampling /
pre-physics actuation.  Every function cites the reference file:line it follows
(paths relative to /root/reference/embodied_pose unless prefixed).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product path (vid2player3d_b200/) never does: it fails loudly when
the CUDA extension is missing.

Pinning: this port is checked (tests/test_oracle_golden.py) against fixtures produced by
EXECUTING the reference's own functions in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz).  The physics half has no reference
source (closed-source Isaac Gym / PhysX): see oracle/physics_ref.c ("parity unpinned").

All functions take/return numpy arrays; `dtype` follows the inputs (float32 mirrors the
reference, float64 gives a tighter bound for kernel tests).  Quaternions are xyzw.
"""
import numpy as np

BASE_ROT_CONJ = np.array([-0.5, -0.5, -0.5, 0.5])  # conj([.5,.5,.5,.5]) humanoid_smpl_im.py:766-770


# --------------------------------------------------------------------------- quaternion helpers
def quat_mul(a, b):
    """isaacgym.torch_utils.quat_mul (xyzw Hamilton product); cross-checked against
    poselib/poselib/core/rotation3d.py:15."""
    x1, y1, z1, w1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    x2, y2, z2, w2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2
    x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2
    y = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2
    z = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2
    return np.stack([x, y, z, w], axis=-1)


def quat_conjugate(a):
    return np.concatenate([-a[..., :3], a[..., 3:]], axis=-1)


def normalize(x, eps=1e-9):
    n = np.linalg.norm(x, axis=-1, keepdims=True)
    return x / np.maximum(n, eps)


def normalize_angle(x):
    return np.arctan2(np.sin(x), np.cos(x))


def quat_from_angle_axis(angle, axis):
    theta = (angle / 2)[..., None]
    xyz = normalize(axis) * np.sin(theta)
    w = np.cos(theta)
    return normalize(np.concatenate([xyz, w], axis=-1))


def my_quat_rotate(q, v):
    """utils/torch_utils.py:70-79"""
    q_w = q[..., 3:4]
    q_vec = q[..., :3]
    a = v * (2.0 * q_w ** 2 - 1.0)
    b = np.cross(q_vec, v) * q_w * 2.0
    c = q_vec * np.sum(q_vec * v, axis=-1, keepdims=True) * 2.0
    return a + b + c


def quat_to_angle_axis(q):
    """utils/torch_utils.py:82-102"""
    min_theta = 1e-5
    with np.errstate(invalid="ignore", divide="ignore"):
        sin_theta = np.sqrt(1 - q[..., 3] * q[..., 3])
        angle = 2 * np.arccos(q[..., 3])
        angle = normalize_angle(angle)
        axis = q[..., 0:3] / sin_theta[..., None]
    mask = np.abs(sin_theta) > min_theta
    default_axis = np.zeros_like(axis)
    default_axis[..., -1] = 1
    angle = np.where(mask, angle, np.zeros_like(angle))
    axis = np.where(mask[..., None], axis, default_axis)
    return angle, axis


def quat_to_exp_map(q):
    """utils/torch_utils.py:113-120"""
    angle, axis = quat_to_angle_axis(q)
    return angle[..., None] * axis


def exp_map_to_angle_axis(exp_map):
    """utils/torch_utils.py:143-160"""
    min_theta = 1e-5
    angle = np.linalg.norm(exp_map, axis=-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = exp_map / angle[..., None]
    angle = normalize_angle(angle)
    default_axis = np.zeros_like(exp_map)
    default_axis[..., -1] = 1
    mask = np.abs(angle) > min_theta
    angle = np.where(mask, angle, np.zeros_like(angle))
    axis = np.where(mask[..., None], axis, default_axis)
    return angle, axis


def exp_map_to_quat(exp_map):
    """utils/torch_utils.py:162-166"""
    angle, axis = exp_map_to_angle_axis(exp_map)
    return quat_from_angle_axis(angle, axis)


def quat_to_tan_norm(q):
    """utils/torch_utils.py:122-134"""
    ref_tan = np.zeros_like(q[..., 0:3])
    ref_tan[..., 0] = 1
    tan = my_quat_rotate(q, ref_tan)
    ref_norm = np.zeros_like(q[..., 0:3])
    ref_norm[..., -1] = 1
    norm = my_quat_rotate(q, ref_norm)
    return np.concatenate([tan, norm], axis=-1)


def slerp(q0, q1, t):
    """utils/torch_utils.py:168-190 (t broadcastable to [...,1])"""
    cos_half_theta = np.sum(q0 * q1, axis=-1)
    neg_mask = cos_half_theta < 0
    q1 = np.where(neg_mask[..., None], -q1, q1)
    cos_half_theta = np.abs(cos_half_theta)[..., None]
    with np.errstate(invalid="ignore", divide="ignore"):
        half_theta = np.arccos(cos_half_theta)
        sin_half_theta = np.sqrt(1.0 - cos_half_theta * cos_half_theta)
        ratioA = np.sin((1 - t) * half_theta) / sin_half_theta
        ratioB = np.sin(t * half_theta) / sin_half_theta
        new_q = ratioA * q0 + ratioB * q1
    new_q = np.where(np.abs(sin_half_theta) < 0.001, 0.5 * q0 + 0.5 * q1, new_q)
    new_q = np.where(np.abs(cos_half_theta) >= 1, q0, new_q)
    return new_q


def calc_heading(q):
    """utils/torch_utils.py:192-203"""
    ref_dir = np.zeros_like(q[..., 0:3])
    ref_dir[..., 0] = 1
    rot_dir = my_quat_rotate(q, ref_dir)
    return np.arctan2(rot_dir[..., 1], rot_dir[..., 0])


def _z_axis_like(q):
    axis = np.zeros_like(q[..., 0:3])
    axis[..., 2] = 1
    return axis


def calc_heading_quat(q):
    """utils/torch_utils.py:205-216"""
    return quat_from_angle_axis(calc_heading(q), _z_axis_like(q))


def calc_heading_quat_inv_with_heading(q):
    """utils/torch_utils.py:232-243"""
    heading = calc_heading(q)
    return quat_from_angle_axis(-heading, _z_axis_like(q)), heading


def remove_base_rot(quat):
    """env/tasks/humanoid_smpl_im.py:766-770"""
    return quat_mul(quat, np.broadcast_to(BASE_ROT_CONJ.astype(quat.dtype), quat.shape))


def heading_to_vec(h):
    """utils/torch_transform.py:188-191"""
    return np.stack([np.cos(h), np.sin(h)], axis=-1)


# --------------------------------------------------------------------------- obs / reward / reset
def dof_to_obs(pose):
    """env/tasks/humanoid_smpl.py:604-635, every joint spherical (dof_size == 3)."""
    n = pose.shape[0]
    q = exp_map_to_quat(pose.reshape(n, -1, 3))
    return quat_to_tan_norm(q).reshape(n, -1)


def compute_humanoid_observations_imitation(body_pos, body_rot, target_pos, target_rot, dof_pos, dof_vel,
                                            target_dof_pos, body_vel, body_ang_vel, motion_bodies,
                                            local_root_obs=True, root_height_obs=True):
    """env/tasks/humanoid_smpl_im.py:773-850 (identical copies: models/im_network_builder.py:262-338,
    vid2player/env/tasks/humanoid_smpl_im_mvae.py:1046-1132).  Returns [N,734] for 24 bodies.
    Quirk kept: with local_root_obs the root slot is tan-norm of the DE-BASED but NOT de-headed
    root rotation (:806-809)."""
    N, B = body_pos.shape[:2]
    root_pos = body_pos[:, 0, :]
    root_rot = body_rot[:, 0, :]
    root_h = root_pos[:, 2:3]
    root_rot = remove_base_rot(root_rot)
    heading_rot, heading = calc_heading_quat_inv_with_heading(root_rot)
    root_h_obs = root_h if root_height_obs else np.zeros_like(root_h)
    hr = np.broadcast_to(heading_rot[:, None, :], (N, B, 4))

    local_body_pos = my_quat_rotate(hr, body_pos - root_pos[:, None, :]).reshape(N, B * 3)[:, 3:]
    local_body_rot_obs = quat_to_tan_norm(quat_mul(hr, body_rot)).reshape(N, B * 6)
    if local_root_obs:
        local_body_rot_obs = local_body_rot_obs.copy()
        local_body_rot_obs[:, 0:6] = quat_to_tan_norm(root_rot)
    local_body_vel = my_quat_rotate(hr, body_vel).reshape(N, B * 3)
    local_body_ang_vel = my_quat_rotate(hr, body_ang_vel).reshape(N, B * 3)

    target_root_pos = target_pos[:, 0, :]
    target_root_rot = remove_base_rot(target_rot[:, 0, :])
    target_rel_root_h = root_h - target_root_pos[:, 2:3]
    _, target_heading = calc_heading_quat_inv_with_heading(target_root_rot)
    target_rel_root_rot_obs = quat_to_tan_norm(quat_mul(target_root_rot, quat_conjugate(root_rot)))
    target_rel_2d_pos = my_quat_rotate(heading_rot, target_root_pos - root_pos)[:, :2]
    target_rel_heading_vec = heading_to_vec(target_heading - heading)
    target_rel_dof_pos = target_dof_pos - dof_pos
    target_rel_body_pos = my_quat_rotate(hr, target_pos - body_pos).reshape(N, B * 3)
    target_rel_body_rot_obs = quat_to_tan_norm(quat_mul(quat_conjugate(body_rot), target_rot)).reshape(N, B * 6)

    return np.concatenate((root_h_obs, local_body_pos, local_body_rot_obs, local_body_vel, local_body_ang_vel,
                           dof_vel, target_rel_root_h, target_rel_root_rot_obs, target_rel_2d_pos,
                           target_rel_heading_vec, target_rel_dof_pos, target_rel_body_pos,
                           target_rel_body_rot_obs, motion_bodies), axis=-1)


def compute_humanoid_observations_imitation_jpos(body_pos, body_rot, target_pos, target_rot, dof_pos, dof_vel,
                                                 target_dof_pos, body_vel, body_ang_vel, motion_bodies,
                                                 local_root_obs=True, root_height_obs=True):
    """env/tasks/humanoid_smpl_im.py:853-915 (obs_type 'joint_pos'): the same features without the rotation / dof targets"""
    full = compute_humanoid_observations_imitation(body_pos, body_rot, target_pos, target_rot, dof_pos, dof_vel, target_dof_pos,
                                                   body_vel, body_ang_vel, motion_bodies, local_root_obs, root_height_obs)
    B, D = body_pos.shape[1], dof_pos.shape[1]
    o = 1 + (B - 1) * 3 + B * 6 + B * 6 + D          # end of dof_vel
    rel_h = full[:, o:o + 1]
    rel_2d = full[:, o + 7:o + 9]
    rel_body = full[:, o + 11 + D:o + 11 + D + B * 3]
    return np.concatenate([full[:, :o], rel_h, rel_2d, rel_body, motion_bodies], axis=-1)


def compute_humanoid_obs_raw(body_pos, body_rot, dof_pos, dof_vel, body_vel, body_ang_vel, motion_bodies):
    """env/tasks/humanoid_smpl_im.py:653-668 with obs_names of :198 -> obs_buf[N,461]."""
    N = body_pos.shape[0]
    return np.concatenate([body_pos.reshape(N, -1), body_rot.reshape(N, -1), dof_pos, dof_vel,
                           body_vel.reshape(N, -1), body_ang_vel.reshape(N, -1), motion_bodies], axis=-1)


DEFAULT_REWARD_SPECS = {'k_dof': 60, 'k_vel': 0.2, 'k_pos': 100, 'k_rot': 40,
                        'w_dof': 0.6, 'w_vel': 0.1, 'w_pos': 0.2, 'w_rot': 0.1}  # humanoid_smpl_im.py:682


def compute_humanoid_reward(body_pos, body_rot, target_pos, target_rot, dof_pos, dof_vel, target_dof_pos,
                            target_dof_vel, body_pos_weights, specs=None):
    """env/tasks/humanoid_smpl_im.py:918-953; returns (reward[N], sub_rewards[N,4])."""
    s = dict(DEFAULT_REWARD_SPECS)
    s.update(specs or {})
    diff_dof_obs = dof_to_obs(dof_pos) - dof_to_obs(target_dof_pos)
    dof_reward = np.exp(-s['k_dof'] * (diff_dof_obs ** 2).mean(axis=-1))
    vel_reward = np.exp(-s['k_vel'] * ((target_dof_vel - dof_vel) ** 2).mean(axis=-1))
    diff_body_pos = (target_pos - body_pos) * body_pos_weights[:, None]
    body_pos_reward = np.exp(-s['k_pos'] * (diff_body_pos ** 2).mean(axis=-1).mean(axis=-1))
    diff_body_rot = quat_mul(target_rot, quat_conjugate(body_rot))
    ang = quat_to_angle_axis(diff_body_rot)[0]
    body_rot_reward = np.exp(-s['k_rot'] * (ang ** 2).mean(axis=-1))
    reward = s['w_dof'] * dof_reward + s['w_vel'] * vel_reward + s['w_pos'] * body_pos_reward + s['w_rot'] * body_rot_reward
    sub = np.stack([dof_reward, vel_reward, body_pos_reward, body_rot_reward], axis=-1)
    return reward.astype(body_pos.dtype), sub.astype(body_pos.dtype)


def compute_reward_caller(rew, sub, reset_buf):
    """env/tasks/humanoid_smpl_im.py:688-691: rows already flagged for reset keep zero reward."""
    m = reset_buf == 1
    rew = np.where(m, 0, rew).astype(rew.dtype)
    sub = np.where(m[:, None], 0, sub).astype(sub.dtype)
    return rew, sub


def compute_humanoid_reset(reset_buf, progress_buf, contact_body_ids, rigid_body_pos, max_episode_length,
                           enable_early_termination, termination_heights, cur_ref_motion_times, ref_motion_lengths):
    """env/tasks/humanoid_smpl_im.py:956-987"""
    terminated = np.zeros_like(reset_buf)
    if enable_early_termination:
        fall_height = rigid_body_pos[..., 2] < termination_heights
        fall_height[:, contact_body_ids] = False
        has_fallen = np.any(fall_height, axis=-1) & (progress_buf > 1)
        terminated = np.where(has_fallen, np.ones_like(reset_buf), terminated)
    reset_cond = (progress_buf >= max_episode_length - 1) | (cur_ref_motion_times >= ref_motion_lengths)
    reset = np.where(reset_cond, np.ones_like(reset_buf), terminated)
    return reset, terminated


def compute_reset_caller(old_reset, old_terminate, reset, terminated):
    """env/tasks/humanoid_smpl_im.py:728-739: reset/terminate are sticky until reset()."""
    m = old_reset == 1
    return np.where(m, 1, reset), np.where(m, old_terminate, terminated)


# --------------------------------------------------------------------------- MoCap buffer
def calc_frame_blend(time, length, num_frames, dt):
    """utils/motion_lib.py:427-436 (float32 arithmetic like the reference)."""
    phase = np.clip(time / length, 0.0, 1.0)
    frame_idx0 = (phase * (num_frames - 1).astype(time.dtype)).astype(np.int64)
    frame_idx1 = np.minimum(frame_idx0 + 1, num_frames - 1)
    blend = (time - frame_idx0.astype(time.dtype) * dt) / dt
    return frame_idx0, frame_idx1, blend


def get_motion_state(ml, motion_ids, motion_times, adjust_height=True, ground_tolerance=0.0):
    """utils/motion_lib.py:164-266 with return_rigid_body=True.
    `ml` is a dict of flat arrays: gts[F,B,3] grs[F,B,4] lrs[F,B,4] grvs[F,3] gravs[F,3] dvs[F,D]
    motion_lengths[M] num_frames[M] motion_dt[M] length_starts[M] min_verts_h[M] key_body_ids[K]
    dof_body_ids[J]."""
    motion_len = ml["motion_lengths"][motion_ids]
    num_frames = ml["num_frames"][motion_ids]
    dt = ml["motion_dt"][motion_ids]
    i0, i1, blend = calc_frame_blend(motion_times, motion_len, num_frames, dt)
    f0 = i0 + ml["length_starts"][motion_ids]
    f1 = i1 + ml["length_starts"][motion_ids]
    blend = blend[:, None]
    bexp = blend[:, :, None]
    root_pos = (1.0 - blend) * ml["gts"][f0, 0] + blend * ml["gts"][f1, 0]
    root_rot = slerp(ml["grs"][f0, 0], ml["grs"][f1, 0], blend)
    kb = ml["key_body_ids"]
    key_pos = (1.0 - bexp) * ml["gts"][f0][:, kb] + bexp * ml["gts"][f1][:, kb]
    local_rot = slerp(ml["lrs"][f0], ml["lrs"][f1], bexp)
    dof_pos = quat_to_exp_map(local_rot[:, ml["dof_body_ids"]]).reshape(len(motion_ids), -1)  # :460-488
    root_vel = ml["grvs"][f0]
    root_ang_vel = ml["gravs"][f0]
    dof_vel = ml["dvs"][f0]
    rb_pos = (1.0 - bexp) * ml["gts"][f0] + bexp * ml["gts"][f1]
    rb_rot = slerp(ml["grs"][f0], ml["grs"][f1], bexp)
    if adjust_height:
        min_vh = ml["min_verts_h"][motion_ids] - ground_tolerance
        root_pos = root_pos.copy()
        root_pos[:, 2] -= min_vh
        key_pos[..., 2] -= min_vh[:, None]
        rb_pos[..., 2] -= min_vh[:, None]
    dt_ = ml["gts"].dtype
    return tuple(x.astype(dt_) for x in (root_pos, root_rot, dof_pos, root_vel, root_ang_vel, dof_vel, key_pos, rb_pos, rb_rot))


# --------------------------------------------------------------------------- pre-physics actuation
def init_context(ml, motion_ids, motion_times, dt, context_length=32, context_padding=8, mask_body_ids=None):
    """HumanoidSMPLIM._init_context + _transform_target 'mask_joints' (embodied_pose/env/tasks/humanoid_smpl_im.py:530-592):
    the window of context_length + 2 padding MoCap frames starting at t + dt - padding dt ->
    feat [n, P, body_pos | body_rot | dof_pos | body_pos_gt | dof_pos_gt (| joint_conf)], mask [n, P] = t_j <= length + 2 dt."""
    dt = np.float32(dt)
    n, P = len(motion_ids), context_length + 2 * context_padding
    steps = dt * np.arange(-context_padding, context_length + context_padding).astype(np.float32)
    all_t = ((motion_times.astype(np.float32) + dt)[:, None] + steps[None, :]).astype(np.float32)
    all_ids = np.repeat(np.asarray(motion_ids)[:, None], P, 1)
    st = get_motion_state(ml, all_ids.reshape(-1), all_t.reshape(-1))
    dof_pos, rb_pos, rb_rot = st[2], st[7], st[8]
    body_pos = rb_pos.copy()
    parts = [None, rb_rot.reshape(n * P, -1), dof_pos, rb_pos.reshape(n * P, -1), dof_pos]
    if mask_body_ids is not None:
        conf = np.ones(rb_pos.shape[:2], np.float32)
        conf[:, list(mask_body_ids)] = 0.0
        body_pos = body_pos * conf[..., None]
        parts.append(conf)
    parts[0] = body_pos.reshape(n * P, -1)
    feat = np.concatenate(parts, -1).astype(np.float32).reshape(n, P, -1)
    mask = all_t <= (ml["motion_lengths"][motion_ids].astype(np.float32) + np.float32(2) * dt)[:, None]
    return feat, mask


def pre_physics(actions, reset_buf, dof_pos, root_body_rot, num_dof, pd_tar_lim, res_force_scale, res_torque_scale):
    """env/tasks/humanoid_smpl_im.py:125-157 + :391-396.
    Returns (actions_used, pd_tar[N,D], force[N,3], torque[N,3]) - the wrench is for body 0, ENV_SPACE."""
    a = actions.copy()
    a[reset_buf == 1] = 0
    dof_a = a[:, :num_dof]
    pd_tar = np.maximum(np.minimum(dof_a, dof_pos + pd_tar_lim), dof_pos - pd_tar_lim)
    f = a[:, num_dof:num_dof + 3] * res_force_scale
    t = a[:, num_dof + 3:num_dof + 6] * res_torque_scale
    hq = calc_heading_quat(remove_base_rot(root_body_rot))
    return a, pd_tar, my_quat_rotate(hq, f), my_quat_rotate(hq, t)


# --------------------------------------------------------------------------- task state machine
class ImTaskOracle:
    """numpy mirror of HumanoidSMPLIM's per-step logic with the physics left to the caller
    (env/tasks/humanoid_smpl_im.py:125-157 pre_physics_step, :398-418 post_physics_step,
    :594-636 targets, :670-692 reward caller, :724-739 reset caller)."""

    def __init__(self, ml, motion_ids, ref_times, progress, reset_buf, terminate_buf, dt, max_episode_length,
                 termination_heights, contact_body_ids, body_pos_weights, motion_bodies,
                 enable_early_termination=True, pd_tar_lim=0.5 * np.pi, res_scale=31.85, reward_specs=None):
        self.ml, self.motion_ids = ml, motion_ids
        self.ref_times = ref_times.copy()
        self.progress = progress.copy()
        self.reset_buf = reset_buf.copy()
        self.terminate_buf = terminate_buf.copy()
        self.dt = np.float32(dt) if ref_times.dtype == np.float32 else dt
        self.max_len = max_episode_length
        self.term_h, self.contact_ids, self.w = termination_heights, contact_body_ids, body_pos_weights
        self.motion_bodies = motion_bodies
        self.early = enable_early_termination
        self.pd_tar_lim, self.res_scale, self.specs = pd_tar_lim, res_scale, reward_specs
        self.set_targets()

    def set_targets(self):
        (self.t_root_pos, self.t_root_rot, self.t_dof_pos, self.t_root_vel, self.t_root_ang_vel, self.t_dof_vel,
         self.t_key_pos, self.t_rb_pos, self.t_rb_rot) = get_motion_state(self.ml, self.motion_ids, self.ref_times + self.dt)

    def pre_physics(self, actions, dof_pos, root_body_rot):
        out = pre_physics(actions, self.reset_buf, dof_pos, root_body_rot, dof_pos.shape[1], self.pd_tar_lim,
                          self.res_scale, self.res_scale)
        self.p_dof_pos, self.p_dof_vel = self.t_dof_pos.copy(), self.t_dof_vel.copy()
        self.p_rb_pos, self.p_rb_rot = self.t_rb_pos.copy(), self.t_rb_rot.copy()
        return out

    def post_physics(self, rbs, dofs):
        """rbs[N,B,13], dofs[N,D,2]: the simulated state after the control step."""
        self.progress = self.progress + 1
        self.ref_times = self.ref_times + self.dt
        self.set_targets()
        bp, br, bv, bw = rbs[..., 0:3], rbs[..., 3:7], rbs[..., 7:10], rbs[..., 10:13]
        dp, dv = dofs[..., 0], dofs[..., 1]
        obs = compute_humanoid_obs_raw(bp, br, dp, dv, bv, bw, self.motion_bodies)
        rew, sub = compute_humanoid_reward(bp, br, self.p_rb_pos, self.p_rb_rot, dp, dv, self.p_dof_pos,
                                           self.p_dof_vel, self.w, self.specs)
        rew, sub = compute_reward_caller(rew, sub, self.reset_buf)
        lens = self.ml["motion_lengths"][self.motion_ids]
        reset, term = compute_humanoid_reset(self.reset_buf, self.progress, self.contact_ids, bp, self.max_len,
                                             self.early, self.term_h, self.ref_times, lens)
        self.reset_buf, self.terminate_buf = compute_reset_caller(self.reset_buf, self.terminate_buf, reset, term)
        return obs, rew, sub
