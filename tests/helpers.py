"""Shared builders for the tests (the config dicts live in vid2player3d_b200/configs.py)."""
import numpy as np

from vid2player3d_b200.configs import SIM_PARAMS, im_cfg, v2p_cfg, v2p_dual_cfg  # noqa: F401


def lib_dict(flat, model, key_names=("R_Ankle", "L_Ankle", "L_Hand", "R_Hand")):
    names = [str(x) for x in model["body_names"]]
    return flat.as_dict([names.index(k) for k in key_names], model["dof_body_ids"])


def rand_quat(rng, *shape):
    q = rng.normal(size=shape + (4,))
    return q / np.linalg.norm(q, axis=-1, keepdims=True)
