"""Hot-path functions of the reference's volsdf/utils/rend_util.py, backed by the CUDA ray kernels.

get_camera_params (rend_util.py:60-95) and get_sphere_intersections (rend_util.py:200-216) keep the
reference's signatures and shapes.  Image IO / pose decomposition helpers of that file are caller-side
host code and are not part of this package.
"""
import torch

from .. import functional as F


def get_camera_params(uv, pose, intrinsics):
    """uv (B,N,2), pose (B,4,4), intrinsics (B,4,4) -> ray_dirs (B,N,3), cam_loc (B,3)."""
    if pose.shape[1] == 7:
        raise NotImplementedError('quaternion poses are unused by the reference datasets (rend_util.py:65-70)')
    dirs = []
    for b in range(uv.shape[0]):
        d, _, _ = F.raygen(uv[b], pose[b], intrinsics[b])
        dirs.append(d)
    return torch.stack(dirs, 0), pose[:, :3, 3]


def get_sphere_intersections(cam_loc, ray_directions, r=1.0):
    """(R,3),(R,3) -> (R,2) near/far distances.  The reference prints 'BOUNDING SPHERE PROBLEM!' and
    exit()s when a ray misses the sphere (rend_util.py:209-211); here that raises instead."""
    nf, bad = F.sphere_intersections(cam_loc, ray_directions, r)
    if torch.cuda.is_current_stream_capturing():
        # a CUDA-graph capture cannot read the flag back; GraphedTrainStep runs (and checks) an eager step with the same
        # static ray buffers before it captures
        return nf
    if int(bad.item()) != 0:
        raise RuntimeError('BOUNDING SPHERE PROBLEM!')
    return nf
