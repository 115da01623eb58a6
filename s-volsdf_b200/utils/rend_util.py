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


# Sticky device-side record of rays that missed the bounding sphere (per device).  Graph replays cannot read a flag back
# per call, and a host read per training step is a synchronisation the reference's loop does not have either (its check
# is a host sync too, but it exit()s): the flag accumulates on the device, `check_bounding_sphere()` reads it.
_BAD = {}


def get_sphere_intersections(cam_loc, ray_directions, r=1.0, _sync=None):
    """(R,3),(R,3) -> (R,2) near/far distances.  The reference prints 'BOUNDING SPHERE PROBLEM!' and exit()s when a ray
    misses the sphere (rend_util.py:209-211).  Here the condition is detected by the kernel; outside CUDA-graph capture
    and outside autograd-recorded (training) calls it raises at once, in training / under capture it is accumulated into
    a per-device flag that `check_bounding_sphere()` (or the next eval call) reports — no host sync in the train step."""
    nf, bad = F.sphere_intersections(cam_loc, ray_directions, r)
    dev = cam_loc.device
    acc = _BAD.get(dev)
    if acc is None:
        acc = _BAD[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    acc.add_(bad)      # in-place on a persistent buffer: also valid inside a captured graph (replays keep accumulating)
    sync = (not torch.cuda.is_current_stream_capturing() and not torch.is_grad_enabled()) if _sync is None else _sync
    if sync:
        check_bounding_sphere(dev)
    return nf


def check_bounding_sphere(device=None):
    """Raises if any ray seen since the last check missed the bounding sphere (one host read)."""
    for dev, acc in list(_BAD.items()):
        if device is not None and dev != device:
            continue
        if int(acc.item()) != 0:
            acc.zero_()
            raise RuntimeError('BOUNDING SPHERE PROBLEM!')
