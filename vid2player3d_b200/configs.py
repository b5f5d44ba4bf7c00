"""Workload configurations of the BASELINE configs as plain dicts (they mirror the reference's yaml files:
embodied_pose/cfg/amass_im.yaml, vid2player/cfg/controller/federer.yaml, federer_djokovic.yaml).  Used by bench.py, tools/ and
the tests."""


def im_cfg(num_envs, motion_lib, episode_length=300, asset="mjcf/smpl_mesh_humanoid_amass_v1.xml", **env_over):
    env = dict(numEnvs=num_envs, envSpacing=5, episodeLength=episode_length, enableDebugVis=False, pdControl=True,
               powerScale=1.0, controlFrequencyInv=2, stateInit="Hybrid", hybridInitProb=1.0, numAMPObsSteps=10,
               localRootObs=True, keyBodies=["R_Ankle", "L_Ankle", "L_Hand", "R_Hand"], contactBodies=["R_Ankle", "L_Ankle"],
               terminationBodyHeight=-0.5, terminationHeadHeight=1.0, enableEarlyTermination=True, motion_lib=motion_lib,
               residual_force_scale=31.85, context_length=32, context_padding=8,
               asset=dict(assetRoot="embodied_pose/data/assets", assetFileName=asset),
               plane=dict(staticFriction=1.0, dynamicFriction=1.0, restitution=0.0))
    env.update(env_over)
    return dict(name="HumanoidSMPLIM", env=env, sim=dict(substeps=2))


SIM_PARAMS = dict(dt=1.0 / 60.0, substeps=2)


def v2p_cfg(num_envs, substeps=6, reward_type="return_w_estimate", early_termination=False, **v2p_over):
    """mirrors vid2player/cfg/controller/federer.yaml (single player)"""
    v2p = dict(player="federer", grip="eastern", court_min=[-5, -16], court_max=[5, -10], racket_friction=0.8, ball_friction=0.2,
               restitution=0.9, spin_scale=5, reward_weights={'pos': 0.1, 'ball_pos': 0.9},
               reward_scales={'pos': 50, 'phase': 10, 'bounce_pos': 1, 'bounce_time': 0.5}, reward_type=reward_type,
               obs_ball_traj_length=10, use_history_ball_obs=False, use_random_ball_target="continuous", vae_action_scale=1.5,
               add_residual_dof="euler", residual_dof_scale=0.4, reset_reaction_nframes=70, random_walk_in_recovery=True)
    v2p.update(v2p_over)
    env = dict(numEnvs=num_envs, episodeLength=300, enableEarlyTermination=early_termination, is_train=True,
               physics=dict(assetFileName="smpl_mesh_humanoid_federer.xml", substeps=substeps, residual_force_scale=31.85, plane_restitution=0.5),
               vid2player=v2p)
    return dict(name="PhysicsMVAEController", env=env, seed=10)


def v2p_dual_cfg(num_envs, assets=("smpl_mesh_humanoid_federer.xml", "smpl_mesh_humanoid_djokovic.xml"), players=("federer", "djokovic"),
                 **v2p_over):
    """mirrors vid2player/cfg/controller/federer_djokovic.yaml (two right-handed players, eastern grips)"""
    cfg = v2p_cfg(num_envs, use_random_ball_target=True, dual_mode="different", player=list(players), grip=["eastern", "eastern"],
                  righthand=[True, True], fix_head_orientation=True, **v2p_over)
    cfg["name"] = "PhysicsMVAEControllerDual"
    cfg["env"]["physics"]["assetFileName"] = list(assets)
    cfg["env"]["physics"]["name"] = "HumanoidSMPLIMMVAEDual"
    return cfg
