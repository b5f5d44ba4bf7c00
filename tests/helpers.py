"""Shared builders for the tests (the config dicts live in vid2player3d_b200/configs.py)."""
import numpy as np

from vid2player3d_b200.configs import SIM_PARAMS, im_cfg, v2p_cfg, v2p_dual_cfg  # noqa: F401


def lib_dict(flat, model, key_names=("R_Ankle", "L_Ankle", "L_Hand", "R_Hand")):
    names = [str(x) for x in model["body_names"]]
    return flat.as_dict([names.index(k) for k in key_names], model["dof_body_ids"])


def rand_quat(rng, *shape):
    q = rng.normal(size=shape + (4,))
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def mixed_decoder_params(seed=5, frame=288, latent=32, hidden=256, experts=6):
    """parameters of a MixedDecoder (vid2player/motion_vae/model.py:186-235) drawn from numpy: expert weights [E, in, out], biases
    [E, out], gate [(weight [out, in], bias [out])] x 3.  Shared by tests/golden/make_golden_nn.py (fills the reference module with
    them) and the GPU test (fills ours)."""
    rng = np.random.default_rng(seed)
    inp, inter, out = latent + frame, latent + hidden, frame + 2
    ws, bs, gate = [], [], []
    for i, o in ((inp, hidden), (inter, hidden), (inter, out)):
        b = np.sqrt(6.0 / i)
        ws.append(rng.uniform(-b, b, size=(experts, i, o)).astype(np.float32))
        bs.append(rng.uniform(-0.1, 0.1, size=(experts, o)).astype(np.float32))
    for i, o in ((inp, 64), (64, 64), (64, experts)):
        b = 1.0 / np.sqrt(i)
        gate.append((rng.uniform(-b, b, size=(o, i)).astype(np.float32) * (4.0 if o == experts else 1.0),
                     rng.uniform(-b, b, size=(o,)).astype(np.float32)))
    return ws, bs, gate
