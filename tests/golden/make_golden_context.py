"""Golden fixture for the 48-frame MoCap context window: the reference's own `HumanoidSMPLIM._init_context` +
`_transform_target` (embodied_pose/env/tasks/humanoid_smpl_im.py:530-592) EXECUTED on a fake `self`, on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_context.py
Output: tests/golden/init_context.npz (committed).  Cases: no transform_specs (378 columns) and the deterministic transform
`mask_joints` (402 columns, joint_conf appended).  `noisy_joints` / `mask_random_joints` draw from torch's CPU generator and are
checked through their invariants in the GPU test instead.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as G  # noqa: E402  (sets up the reference import harness)


def init_context():
    model, flat, key = G.small_lib()
    ml = G.make_ref_motion_lib(flat, key, model["dof_body_ids"])
    g = torch.Generator().manual_seed(21)
    N = 16
    rec = {}
    for case, specs in (("plain", None), ("mask", {'mask_joints': {'joints': ['L_Knee', 'R_Wrist', 'Head']}})):
        t = G.FakeTask()
        t.cfg = {'env': {} if specs is None else {'transform_specs': specs}}
        t.device = 'cpu'
        t.num_envs = N
        t.dt = 2 * (1.0 / 60.0)
        t.context_length, t.context_padding = 32, 8
        t._motion_lib = ml
        t.ground_tolerance = 0.0
        t.model = None
        t.body_names = [str(x) for x in model["body_names"]]
        t.context_names = ['body_pos', 'body_rot', 'dof_pos', 'body_pos_gt', 'dof_pos_gt'] + (['joint_conf'] if specs else [])
        ids = torch.randint(0, 6, (N,), generator=g)
        times = torch.rand(N, generator=g) * 1.2
        times[:4] = torch.tensor([0.0, 0.3, 5.0, 1e-3])          # before the start / past the end of the clip: clamped frames, mask off
        t._reset_ref_motion_ids = ids
        t._init_context(ids, times)
        rec[f"{case}_ids"], rec[f"{case}_times"] = ids, times
        rec[f"{case}_feat"], rec[f"{case}_mask"] = t.context_feat, t.context_mask
    G.npz("init_context.npz", **{"lib_" + k: getattr(flat, k) for k in flat.FIELDS}, key_body_ids=np.array(key),
          dof_body_ids=model["dof_body_ids"], mask_joints=np.array(['L_Knee', 'R_Wrist', 'Head']), **rec)


if __name__ == "__main__":
    init_context()
