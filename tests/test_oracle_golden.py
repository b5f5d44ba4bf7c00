"""Pin the numpy oracle (oracle/ref_port.py) against fixtures produced by executing the
reference's own functions (tests/golden/make_golden.py).  CPU only."""
import numpy as np

from conftest import golden
from oracle import ref_port as R

TOL = 2e-6  # float32 round-off between torch's op order and numpy's


def lib_from(g):
    ml = {k[4:]: g[k] for k in g if k.startswith("lib_")}
    ml["key_body_ids"] = g["key_body_ids"].astype(np.int64)
    ml["dof_body_ids"] = g["dof_body_ids"].astype(np.int64)
    return ml


def close(a, b, tol=TOL):
    np.testing.assert_allclose(a, b, rtol=0, atol=tol)


def test_primitives():
    g = golden("primitives.npz")
    q, q2, v, t, e = g["q"], g["q2"], g["v"], g["t"], g["e"]
    close(R.quat_mul(q, q2), g["quat_mul"])
    close(R.quat_conjugate(q), g["quat_conjugate"])
    close(R.my_quat_rotate(q, v), g["rotate"])
    ang, axis = R.quat_to_angle_axis(q)
    close(ang, g["angle"], 1e-5)
    close(axis, g["axis"], 2e-4)  # axis of near-identity quats is ill-conditioned in float32
    close(R.quat_to_exp_map(q), g["exp_map"], 1e-5)
    close(R.quat_to_tan_norm(q), g["tan_norm"])
    close(R.exp_map_to_quat(e), g["exp_to_quat"])
    close(R.slerp(q, q2, t), g["slerp"], 1e-5)
    close(R.calc_heading(q), g["heading"])
    hq, h = R.calc_heading_quat_inv_with_heading(q)
    close(hq, g["heading_q_inv"])
    close(R.calc_heading_quat(q), g["heading_q"])
    close(R.remove_base_rot(q), g["remove_base"])
    close(R.heading_to_vec(g["heading"]), g["heading_vec"])
    close(R.normalize_angle(v[:, 0] * 3), g["normalize_angle"])


def test_obs_imitation():
    g = golden("obs_imitation.npz")
    args = [g[k] for k in ("body_pos", "body_rot", "target_pos", "target_rot", "dof_pos", "dof_vel", "target_dof_pos",
                           "body_vel", "body_ang_vel", "motion_bodies")]
    close(R.compute_humanoid_observations_imitation(*args, True, True), g["obs"], 1e-5)
    close(R.compute_humanoid_observations_imitation(*args, False, False), g["obs_nolocal_noheight"], 1e-5)
    close(R.compute_humanoid_observations_imitation_jpos(*args, True, True), g["obs_jpos"], 1e-5)
    assert g["obs_jpos"].shape[1] == 513
    a64 = [a.astype(np.float64) for a in args]
    close(R.compute_humanoid_observations_imitation(*a64, True, True), g["obs_f64"], 1e-12)
    # float32 reference vs float64 oracle: the tolerance the CUDA kernel is held to
    close(R.compute_humanoid_observations_imitation(*a64, True, True), g["obs"], 1e-5)


def test_dof_reward_reset():
    g = golden("dof_reward_reset.npz")
    close(R.dof_to_obs(g["dof_pos"]), g["dof_obs"])
    rew, sub = R.compute_humanoid_reward(g["body_pos"], g["body_rot"], g["target_pos"], g["target_rot"], g["dof_pos"],
                                         g["dof_vel"], g["target_dof_pos"], g["target_dof_vel"], g["weights"])
    close(rew, g["reward"], 1e-5)
    close(sub, g["sub_rewards"], 1e-5)
    assert str(g["names"]) == "dof_reward,vel_reward,body_pos_reward,body_rot_reward"
    N = len(g["progress"])
    for early, kr, kt in ((True, "reset", "terminated"), (False, "reset_noearly", "terminated_noearly")):
        reset, term = R.compute_humanoid_reset(np.zeros(N, np.int64), g["progress"], g["contact_ids"], g["rb_pos"],
                                               300.0, early, g["heights"], g["times"], g["lens"])
        assert np.array_equal(reset, g[kr]) and np.array_equal(term, g[kt])
    assert g["terminated"].sum() > 0 and g["reset"].sum() > g["terminated"].sum()


def test_motion_state():
    g = golden("motion_state.npz")
    ml = lib_from(g)
    res = R.get_motion_state(ml, g["motion_ids"], g["motion_times"])
    for name, r in zip(("root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "key_pos", "rb_pos",
                        "rb_rot"), res):
        close(r, g[name], 2e-5)


def test_im_step_state_machine():
    g = golden("im_step.npz")
    ml = lib_from(g)
    N = len(g["motion_ids"])
    o = R.ImTaskOracle(ml, g["motion_ids"], g["init_ref_times"], g["init_progress"], g["init_reset"],
                       g["init_terminate"], 2 * (1.0 / 60.0), int(g["max_episode_length"]), g["termination_heights"],
                       g["contact_body_ids"], np.ones(24, np.float32), ml["motion_bodies"][g["motion_ids"]])
    close(o.t_dof_pos, g["init_target_dof_pos"], 2e-5)
    rbs, dofs = g["init_rbs"], g["init_dofs"]
    saw_reset = 0
    for s in range(int(g["steps"])):
        a, pd, f, t = o.pre_physics(g[f"actions_{s}"], dofs[..., 0], rbs[:, 0, 3:7])
        close(pd, g[f"pd_tar_{s}"])
        close(f, g[f"force_{s}"], 1e-4)  # |f| ~ 30: 3e-6 relative
        close(t, g[f"torque_{s}"], 1e-4)
        rbs, dofs = g[f"rbs_{s}"], g[f"dofs_{s}"]
        obs, rew, sub = o.post_physics(rbs, dofs)
        close(obs, g[f"obs_{s}"])
        close(rew, g[f"rew_{s}"], 1e-5)
        close(sub, g[f"sub_{s}"], 1e-5)
        assert np.array_equal(o.reset_buf, g[f"reset_{s}"])
        assert np.array_equal(o.terminate_buf, g[f"terminate_{s}"])
        assert np.array_equal(o.progress, g[f"progress_{s}"])
        close(o.ref_times, g[f"ref_times_{s}"], 1e-6)
        close(o.t_dof_pos, g[f"target_dof_pos_{s}"], 2e-5)
        close(o.t_rb_pos, g[f"target_rb_pos_{s}"], 2e-5)
        close(o.t_rb_rot, g[f"target_rb_rot_{s}"], 2e-5)
        close(o.t_key_pos, g[f"target_key_pos_{s}"], 2e-5)
        saw_reset = int(o.reset_buf.sum())
    assert saw_reset >= 3  # sticky flag + fall + episode end all exercised


def test_init_context():
    """oracle init_context vs the reference's own _init_context / _transform_target run on a fake self (make_golden_context.py)"""
    g = golden("init_context.npz")
    ml = lib_from(g)
    dt = np.float32(2) * np.float32(1 / 60)
    feat, mask = R.init_context(ml, g["plain_ids"], g["plain_times"], dt)
    assert feat.shape == g["plain_feat"].shape == (16, 48, 378)
    close(feat, g["plain_feat"], 1e-5)
    assert np.array_equal(mask, g["plain_mask"]) and mask.any() and not mask.all()
    names = ['Pelvis', 'L_Hip', 'L_Knee', 'L_Ankle', 'L_Toe', 'R_Hip', 'R_Knee', 'R_Ankle', 'R_Toe', 'Torso', 'Spine', 'Chest', 'Neck', 'Head',
             'L_Thorax', 'L_Shoulder', 'L_Elbow', 'L_Wrist', 'L_Hand', 'R_Thorax', 'R_Shoulder', 'R_Elbow', 'R_Wrist', 'R_Hand']
    ids = [names.index(str(j)) for j in g["mask_joints"]]
    feat, mask = R.init_context(ml, g["mask_ids"], g["mask_times"], dt, mask_body_ids=ids)
    assert feat.shape == g["mask_feat"].shape == (16, 48, 402)
    close(feat, g["mask_feat"], 1e-5)
    assert np.array_equal(mask, g["mask_mask"])
