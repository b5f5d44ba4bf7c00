"""Generate the golden fixtures by EXECUTING the reference's own PyTorch functions on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Outputs tests/golden/*.npz (committed).  The reference modules are imported unchanged
through the isaacgym shim (tests/golden/_refenv.py); nothing is copied from them.

Fixtures
  primitives.npz     utils/torch_utils.py helpers + isaacgym.torch_utils shims
  obs_imitation.npz  compute_humanoid_observations_imitation (humanoid_smpl_im.py:773-850)
  dof_reward_reset.npz  dof_to_obs, compute_humanoid_reward, compute_humanoid_reset
  motion_state.npz   MotionLib.get_motion_state on a ragged synthetic library (motion_lib.py:164-266)
  im_step.npz        HumanoidSMPLIM.pre_physics_step / post_physics_step driven through the real
                     methods on a fake `self` (physics replaced by injected states) for 4 steps:
                     pins the prev-target / sticky-reset / zero-reward-on-reset state machine.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _refenv  # noqa: E402

_refenv.setup("embodied_pose")

from env.tasks import humanoid_smpl_im as H  # noqa: E402
from env.tasks import humanoid_smpl as HS  # noqa: E402
from utils import torch_utils as TU  # noqa: E402
from utils.torch_transform import heading_to_vec  # noqa: E402
from utils.motion_lib import MotionLib  # noqa: E402
from isaacgym import torch_utils as IG  # noqa: E402

from vid2player3d_b200 import model_compiler, motion_lib as fml  # noqa: E402

torch.set_num_threads(1)


def rq(g, *s):
    q = torch.randn(*s, 4, generator=g)
    return q / q.norm(dim=-1, keepdim=True)


def npz(name, **kw):
    out = {}
    for k, v in kw.items():
        out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items()})


def primitives():
    g = torch.Generator().manual_seed(1)
    n = 256
    q = rq(g, n)
    q2 = rq(g, n)
    q[0] = torch.tensor([0, 0, 0, 1.0])
    q[1] = torch.tensor([0, 0, 0, -1.0])
    q2[2] = q[2]
    q2[3] = -q[3]
    q2[4] = torch.nn.functional.normalize(q[4] + 1e-4 * torch.randn(4, generator=g), dim=-1)
    v = torch.randn(n, 3, generator=g)
    t = torch.rand(n, 1, generator=g)
    e = torch.randn(n, 3, generator=g)
    e[0] = 0
    e[1] = torch.tensor([1e-6, 0, 0])
    e[2] = torch.tensor([0, 3.5, 0])  # > pi: normalize_angle wraps
    ang, axis = TU.quat_to_angle_axis(q)
    hq_inv, heading = TU.calc_heading_quat_inv_with_heading(q)
    npz("primitives.npz", q=q, q2=q2, v=v, t=t, e=e,
        quat_mul=IG.quat_mul(q, q2), quat_conjugate=IG.quat_conjugate(q), rotate=TU.my_quat_rotate(q, v),
        angle=ang, axis=axis, exp_map=TU.quat_to_exp_map(q), tan_norm=TU.quat_to_tan_norm(q),
        exp_to_quat=TU.exp_map_to_quat(e), slerp=TU.slerp(q, q2, t), heading=heading, heading_q_inv=hq_inv,
        heading_q=TU.calc_heading_quat(q), remove_base=H.remove_base_rot(q),
        heading_vec=heading_to_vec(heading), normalize_angle=IG.normalize_angle(v[:, 0] * 3))


def obs_imitation():
    g = torch.Generator().manual_seed(2)
    N, B, D = 64, 24, 69
    inp = dict(body_pos=torch.randn(N, B, 3, generator=g), body_rot=rq(g, N, B),
               target_pos=torch.randn(N, B, 3, generator=g), target_rot=rq(g, N, B),
               dof_pos=torch.randn(N, D, generator=g), dof_vel=3 * torch.randn(N, D, generator=g),
               target_dof_pos=torch.randn(N, D, generator=g), body_vel=torch.randn(N, B, 3, generator=g),
               body_ang_vel=torch.randn(N, B, 3, generator=g), motion_bodies=torch.randn(N, 11, generator=g))
    obs = H.compute_humanoid_observations_imitation(*inp.values(), True, True)
    obs_nl = H.compute_humanoid_observations_imitation(*inp.values(), False, False)
    obs64 = H.compute_humanoid_observations_imitation(*[x.double() for x in inp.values()], True, True)
    obs_jpos = H.compute_humanoid_observations_imitation_jpos(*inp.values(), True, True)
    npz("obs_imitation.npz", **inp, obs=obs, obs_nolocal_noheight=obs_nl, obs_f64=obs64, obs_jpos=obs_jpos)


def dof_reward_reset():
    g = torch.Generator().manual_seed(3)
    N, B, D = 64, 24, 69
    offs = list(range(0, D + 1, 3))
    dof_pos = 0.8 * torch.randn(N, D, generator=g)
    tdof = dof_pos + 0.1 * torch.randn(N, D, generator=g)
    dof_vel = 2 * torch.randn(N, D, generator=g)
    tdof_vel = dof_vel + torch.randn(N, D, generator=g)
    body_pos = torch.randn(N, B, 3, generator=g)
    tpos = body_pos + 0.05 * torch.randn(N, B, 3, generator=g)
    body_rot = rq(g, N, B)
    trot = torch.nn.functional.normalize(body_rot + 0.1 * torch.randn(N, B, 4, generator=g), dim=-1)
    trot[0] = body_rot[0]
    w = torch.ones(B)
    w[[3, 7]] = 2.0
    specs = {'k_dof': 60., 'k_vel': 0.2, 'k_pos': 100., 'k_rot': 40., 'w_dof': 0.6, 'w_vel': 0.1, 'w_pos': 0.2, 'w_rot': 0.1}
    rew, sub, names = H.compute_humanoid_reward(body_pos, body_rot, tpos, trot, dof_pos, dof_vel, tdof, tdof_vel,
                                                torch.zeros(N, B, 3), torch.zeros(N, B, 3), 138, offs, w, specs)
    # reset
    reset_buf = torch.zeros(N, dtype=torch.long)
    progress = torch.randint(0, 300, (N,), generator=g)
    progress[:4] = torch.tensor([0, 1, 2, 299])
    rb = body_pos.clone()
    rb[..., 2] = torch.rand(N, B, generator=g) * 2 - 0.6
    heights = torch.full((B,), -0.5)
    heights[13] = 1.0
    times = torch.rand(N, generator=g) * 10
    lens = torch.full((N,), 9.0)
    contact_ids = torch.tensor([7, 3])
    reset, term = H.compute_humanoid_reset(reset_buf, progress, torch.zeros(N, B, 3), contact_ids, rb, 300.0, True,
                                           heights, times, lens)
    reset_ne, term_ne = H.compute_humanoid_reset(reset_buf, progress, torch.zeros(N, B, 3), contact_ids, rb, 300.0,
                                                 False, heights, times, lens)
    npz("dof_reward_reset.npz", dof_pos=dof_pos, target_dof_pos=tdof, dof_vel=dof_vel, target_dof_vel=tdof_vel,
        body_pos=body_pos, target_pos=tpos, body_rot=body_rot, target_rot=trot, weights=w,
        dof_obs=HS.dof_to_obs(dof_pos, 138, offs), reward=rew, sub_rewards=sub, names=np.array(names),
        progress=progress, rb_pos=rb, heights=heights, times=times, lens=lens, contact_ids=contact_ids,
        reset=reset, terminated=term, reset_noearly=reset_ne, terminated_noearly=term_ne)


def make_ref_motion_lib(flat, key_body_ids, dof_body_ids):
    ml = MotionLib.__new__(MotionLib)
    ml._device = 'cpu'
    ml._dof_body_ids = list(dof_body_ids)
    ml._dof_offsets = list(range(0, 3 * len(dof_body_ids) + 1, 3))
    ml._num_dof = 3 * len(dof_body_ids)
    ml._key_body_ids = torch.tensor(key_body_ids)
    for k in ("gts", "grs", "lrs", "grvs", "gravs", "dvs"):
        setattr(ml, k, torch.from_numpy(getattr(flat, k)))
    ml._motion_lengths = torch.from_numpy(flat.motion_lengths)
    ml._motion_num_frames = torch.from_numpy(flat.num_frames)
    ml._motion_dt = torch.from_numpy(flat.motion_dt)
    ml._motion_min_verts_h = torch.from_numpy(flat.min_verts_h)
    ml._motion_bodies = torch.from_numpy(flat.motion_bodies)
    ml.generate_length_starts()
    assert np.array_equal(ml.length_starts.numpy(), flat.length_starts)
    ml.motion_ids = torch.arange(flat.num_motions())
    return ml


def small_lib():
    model = model_compiler.load_compiled("smpl_mesh_humanoid_amass_v1")
    flat = fml.synthetic(model, num_motions=6, num_frames=40, seed=11, sigma=0.08, ragged=True)
    flat.min_verts_h = np.linspace(-0.02, 0.03, 6).astype(np.float32)
    flat.motion_bodies = np.random.default_rng(5).normal(size=(6, 11)).astype(np.float32)
    names = list(model["body_names"])
    key = [names.index(n) for n in ["R_Ankle", "L_Ankle", "L_Hand", "R_Hand"]]
    return model, flat, key


def motion_state():
    model, flat, key = small_lib()
    ml = make_ref_motion_lib(flat, key, model["dof_body_ids"])
    g = torch.Generator().manual_seed(4)
    n = 96
    ids = torch.randint(0, 6, (n,), generator=g)
    times = torch.rand(n, generator=g) * 1.5
    times[:6] = torch.tensor([0.0, -0.05, 1.0 / 30, 5.0, 0.5, 1e-4])
    times[6] = ml._motion_lengths[ids[6]]
    res = ml.get_motion_state(ids, times, return_rigid_body=True, adjust_height=True, ground_tolerance=0.0)
    names = ("root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "key_pos", "rb_pos", "rb_rot")
    npz("motion_state.npz", **{"lib_" + k: getattr(flat, k) for k in flat.FIELDS}, key_body_ids=np.array(key),
        dof_body_ids=model["dof_body_ids"], motion_ids=ids, motion_times=times, **dict(zip(names, res)))


class FakeTask(H.HumanoidSMPLIM):
    def __init__(self):  # bypass Isaac construction; attributes are set by hand below
        pass


def im_step():
    model, flat, key = small_lib()
    ml = make_ref_motion_lib(flat, key, model["dof_body_ids"])
    g = torch.Generator().manual_seed(6)
    N, B, D = 32, 24, 69
    names = list(model["body_names"])
    t = FakeTask()
    t.cfg = {'env': {}}
    t.device = 'cpu'
    t.num_envs, t.num_bodies, t._num_dof, t.num_dof = N, B, D, D
    t._pd_control = True
    t.pd_tar_lim = 0.5 * np.pi
    t.residual_force_scale = t.residual_torque_scale = 31.85
    t.stiffness = torch.ones(D)
    t.gym = MagicMock()
    t.sim = None
    t.viewer = None
    t.debug_viz = False
    t.dt = 2 * (1.0 / 60.0)
    t._motion_lib = ml
    t.ground_tolerance = 0.0
    t.max_episode_length = 12
    t._enable_early_termination = True
    t._termination_heights = torch.full((B,), -0.5)
    t._termination_heights[names.index("Head")] = 1.0
    t._contact_body_ids = torch.tensor([names.index("R_Ankle"), names.index("L_Ankle")])
    t._dof_obs_size = 138
    t._dof_offsets = list(range(0, D + 1, 3))
    t.body_pos_weights = torch.ones(B)
    t.obs_names = ['body_pos', 'body_rot', 'dof_pos', 'dof_vel', 'body_vel', 'body_ang_vel', 'motion_bodies']
    t._state_reset_happened = False
    t._reset_ref_env_ids = []
    t.extras = {}
    t._sub_rewards = None
    t._sub_rewards_names = None
    t.obs_buf = torch.zeros(N, 461)
    t.rew_buf = torch.zeros(N)
    t.reset_buf = torch.zeros(N, dtype=torch.long)
    t._terminate_buf = torch.zeros(N, dtype=torch.long)
    t.progress_buf = torch.randint(0, 10, (N,), generator=g)
    t._reset_ref_motion_ids = torch.randint(0, 6, (N,), generator=g)
    t._reset_ref_motion_bodies = ml._motion_bodies[t._reset_ref_motion_ids]
    t._cur_ref_motion_times = torch.rand(N, generator=g) * 1.0
    t._contact_forces = torch.zeros(N, B, 3)
    rbs = torch.zeros(N, B, 13)
    t._rigid_body_state = rbs
    t._rigid_body_pos, t._rigid_body_rot = rbs[..., 0:3], rbs[..., 3:7]
    t._rigid_body_vel, t._rigid_body_ang_vel = rbs[..., 7:10], rbs[..., 10:13]
    dofs = torch.zeros(N, D, 2)
    t._dof_pos, t._dof_vel = dofs[..., 0], dofs[..., 1]
    t._set_target_motion_state()          # initial targets (as reset() would)
    t.reset_buf[1] = 1                    # an env already flagged: action zeroed, reward zeroed, sticky
    t._terminate_buf[1] = 1

    def inject_state(step):
        # "physics result": the MoCap pose of the current ref time plus noise
        _, _, dof_pos, _, _, dof_vel, _, rb_pos, rb_rot = ml.get_motion_state(
            t._reset_ref_motion_ids, t._cur_ref_motion_times + t.dt, return_rigid_body=True, adjust_height=True)
        rbs[..., 0:3] = rb_pos + 0.03 * torch.randn(N, B, 3, generator=g)
        rbs[..., 3:7] = torch.nn.functional.normalize(rb_rot + 0.05 * torch.randn(N, B, 4, generator=g), dim=-1)
        rbs[..., 7:13] = torch.randn(N, B, 6, generator=g)
        dofs[..., 0] = dof_pos + 0.05 * torch.randn(N, D, generator=g)
        dofs[..., 1] = dof_vel + 0.5 * torch.randn(N, D, generator=g)
        if step == 1:
            rbs[5, 4, 2] = -0.7        # a body falls below the termination height
        if step == 2:
            rbs[6, 13, 2] = 0.4        # head below 1.0

    inject_state(-1)
    rec = {"init_progress": t.progress_buf.clone(), "init_ref_times": t._cur_ref_motion_times.clone(),
           "motion_ids": t._reset_ref_motion_ids.clone(), "init_reset": t.reset_buf.clone(),
           "init_terminate": t._terminate_buf.clone(), "init_rbs": rbs.clone(), "init_dofs": dofs.clone(),
           "init_target_dof_pos": t._target_dof_pos.clone(), "init_target_rb_pos": t._target_rb_pos.clone(),
           "init_target_rb_rot": t._target_rb_rot.clone(), "init_target_dof_vel": t._target_dof_vel.clone()}
    S = 4
    for s in range(S):
        actions = torch.clamp(torch.randn(N, 75, generator=g), -1, 1)
        if s == 1:
            actions[:, :69] *= 3.0     # exercise the +-0.5pi clamp
        rec[f"actions_{s}"] = actions.clone()
        t.gym.reset_mock()
        t.pre_physics_step(actions.clone())
        pd_tar = t.gym.set_dof_position_target_tensor.call_args[0][1]
        forces, torques, space = t.gym.apply_rigid_body_force_tensors.call_args[0][1:4]
        rec[f"pd_tar_{s}"] = pd_tar.clone()
        rec[f"force_{s}"] = forces[:, 0].clone()
        rec[f"torque_{s}"] = torques[:, 0].clone()
        assert torch.count_nonzero(forces[:, 1:]) == 0
        inject_state(s)
        rec[f"rbs_{s}"] = rbs.clone()
        rec[f"dofs_{s}"] = dofs.clone()
        t.post_physics_step()
        rec[f"obs_{s}"] = t.obs_buf.clone()
        rec[f"rew_{s}"] = t.rew_buf.clone()
        rec[f"sub_{s}"] = t._sub_rewards.clone()
        rec[f"reset_{s}"] = t.reset_buf.clone()
        rec[f"terminate_{s}"] = t._terminate_buf.clone()
        rec[f"progress_{s}"] = t.progress_buf.clone()
        rec[f"ref_times_{s}"] = t._cur_ref_motion_times.clone()
        rec[f"target_dof_pos_{s}"] = t._target_dof_pos.clone()
        rec[f"target_rb_pos_{s}"] = t._target_rb_pos.clone()
        rec[f"target_rb_rot_{s}"] = t._target_rb_rot.clone()
        rec[f"target_root_pos_{s}"] = t._target_root_pos.clone()
        rec[f"target_key_pos_{s}"] = t._target_key_pos.clone()
    npz("im_step.npz", **{"lib_" + k: getattr(flat, k) for k in flat.FIELDS}, key_body_ids=np.array(key),
        dof_body_ids=model["dof_body_ids"], termination_heights=t._termination_heights,
        contact_body_ids=t._contact_body_ids, max_episode_length=np.array(t.max_episode_length), steps=np.array(S),
        sub_names=np.array(t._sub_rewards_names), **rec)


if __name__ == "__main__":
    primitives()
    obs_imitation()
    dof_reward_reset()
    motion_state()
    im_step()
