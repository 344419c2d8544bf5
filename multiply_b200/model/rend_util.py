"""Mirror of /root/reference/code/lib/utils/rend_util.py (the two functions on the hot path)."""
import torch
import torch.nn.functional as F


def get_camera_params_host(uv, pose, intrinsics):
    """Host (CPU) evaluation of get_camera_params (rend_util.py:45-87) used only to build
    ray/box hit lists, which the reference also computes on the host (multiply.py:256).
    Returns (ray_dirs [R,3], cam_loc [R,3])."""
    cam_loc = pose[:, :3, 3]
    b, n, _ = uv.shape
    x = uv[:, :, 0].view(b, -1)
    y = uv[:, :, 1].view(b, -1)
    z = torch.ones((b, n))
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy, sk = intrinsics[:, 0, 2], intrinsics[:, 1, 2], intrinsics[:, 0, 1]
    xl = (x - cx[:, None] + cy[:, None] * sk[:, None] / fy[:, None] - sk[:, None] * y / fy[:, None]) / fx[:, None] * z
    yl = (y - cy[:, None]) / fy[:, None] * z
    pts = torch.stack((xl, yl, z, torch.ones_like(z)), dim=-1).permute(0, 2, 1)
    world = torch.bmm(pose, pts).permute(0, 2, 1)[:, :, :3]
    dirs = F.normalize(world - cam_loc[:, None, :], dim=2)
    return dirs.reshape(-1, 3), cam_loc[:, None, :].expand(-1, n, -1).reshape(-1, 3)
