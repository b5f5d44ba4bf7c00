"""Small autograd-carrying torch helpers used by get_aux_losses (stays PyTorch per north_star)."""
import torch


def angle_axis_to_rot6d(aa):
    """rotation vector [...,3] -> first two columns of the rotation matrix, flattened like the
    reference's utils/torch_transform.py::angle_axis_to_rot6d (rotmat[..., :2] transposed)."""
    angle = aa.norm(dim=-1, keepdim=True)
    axis = aa / angle.clamp_min(1e-8)
    x, y, z = axis.unbind(-1)
    c, s = torch.cos(angle[..., 0]), torch.sin(angle[..., 0])
    C = 1 - c
    R = torch.stack([c + x * x * C, x * y * C - z * s, x * z * C + y * s,
                     y * x * C + z * s, c + y * y * C, y * z * C - x * s,
                     z * x * C - y * s, z * y * C + x * s, c + z * z * C], dim=-1).view(*aa.shape[:-1], 3, 3)
    return R[..., :2].transpose(-1, -2).reshape(*aa.shape[:-1], 6)


def quaternion_wxyz_to_angle_axis(q, eps=1.0e-6):
    """rotation vector of a (w, x, y, z) quaternion, the ceres formula of the reference's utils/konia_transform.py:558-628
    (test-time export of the simulated pose, humanoid_smpl_im_mvae.py:814-820; plain torch, not on the training path)"""
    cos_t, q1, q2, q3 = q.unbind(-1)
    s2 = q1 * q1 + q2 * q2 + q3 * q3
    sin_t = torch.sqrt(s2.clamp_min(eps))
    two_theta = 2.0 * torch.where(cos_t < 0.0, torch.atan2(-sin_t, -cos_t), torch.atan2(sin_t, cos_t))
    k = torch.where(s2 > 0.0, two_theta / sin_t, torch.full_like(sin_t, 2.0))
    return torch.stack([q1 * k, q2 * k, q3 * k], dim=-1)


def get_opponent_env_ids(env_ids):
    """vid2player/utils/common.py:111-115: the paired env of each id (2k <-> 2k+1), sorted"""
    if len(env_ids) == 0:
        return env_ids
    return (env_ids ^ 1).sort().values
