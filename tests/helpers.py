"""Shared builders for the tests (config dicts mirror embodied_pose/cfg/amass_im.yaml)."""
import numpy as np


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


def lib_dict(flat, model, key_names=("R_Ankle", "L_Ankle", "L_Hand", "R_Hand")):
    names = [str(x) for x in model["body_names"]]
    return flat.as_dict([names.index(k) for k in key_names], model["dof_body_ids"])


def rand_quat(rng, *shape):
    q = rng.normal(size=shape + (4,))
    return q / np.linalg.norm(q, axis=-1, keepdims=True)
