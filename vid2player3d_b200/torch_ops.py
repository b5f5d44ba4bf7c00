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


def get_opponent_env_ids(env_ids):
    """vid2player/utils/common.py:111-115: the paired env of each id (2k <-> 2k+1), sorted"""
    if len(env_ids) == 0:
        return env_ids
    return (env_ids ^ 1).sort().values
