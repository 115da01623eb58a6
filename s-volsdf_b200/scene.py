"""Synthetic DTU / BlendedMVS-shaped scenes used by tests, goldens and bench.py.

There is no dataset on the GPU box, so rays come from a DTU-like pinhole camera looking at the
geometric-init sphere (radius ~0.6) from inside the R=3 bounding sphere.  The pixel grid follows the
reference's (x, y) order (volsdf/datasets/scene_dataset.py:227-229).
"""
import numpy as np
import torch


def dtu_camera(width=1600, height=1200):
    s = width / 1600.0
    K = torch.eye(4, dtype=torch.float32)
    K[0, 0] = 2892.33 * s
    K[1, 1] = 2883.18 * s
    K[0, 2] = 823.2 * s
    K[1, 2] = 619.07 * s
    pose = torch.eye(4, dtype=torch.float32)
    pose[2, 3] = -2.5
    return K[None], pose[None]


def bmvs_camera(width=768, height=576):
    K = torch.eye(4, dtype=torch.float32)
    K[0, 0] = 700.0
    K[1, 1] = 700.0
    K[0, 2] = width / 2.0
    K[1, 2] = height / 2.0
    pose = torch.eye(4, dtype=torch.float32)
    pose[2, 3] = -2.0
    return K[None], pose[None]


def random_pixels(n, width, height, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, width, (n,), generator=g)
    y = torch.randint(0, height, (n,), generator=g)
    return torch.stack([x, y], -1).float()[None]


def permuted_pixels(n, width, height, seed=1):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(width * height, generator=g)[:n]
    return torch.stack([idx % width, idx // width], -1).float()[None]


def full_grid(width, height):
    uv = np.mgrid[0:height, 0:width].astype(np.int32)
    uv = torch.from_numpy(np.flip(uv, axis=0).copy()).float()
    return uv.reshape(2, -1).transpose(1, 0)[None].contiguous()


def make_input(kind='dtu', n_rays=1024, width=None, height=None, seed=1, pixels='random'):
    """Returns the reference's model_input dict (CPU tensors): intrinsics (1,4,4), uv (1,R,2), pose (1,4,4)."""
    if kind == 'dtu':
        width, height = width or 1600, height or 1200
        K, pose = dtu_camera(width, height)
    elif kind == 'bmvs':
        width, height = width or 768, height or 576
        K, pose = bmvs_camera(width, height)
    else:
        raise ValueError(kind)
    if pixels == 'random':
        uv = random_pixels(n_rays, width, height, seed)
    elif pixels == 'perm':
        uv = permuted_pixels(n_rays, width, height, seed)
    elif pixels == 'grid':
        uv = full_grid(width, height)
    else:
        raise ValueError(pixels)
    inp = {'intrinsics': K, 'uv': uv, 'pose': pose}
    if kind == 'bmvs':
        inp['near_pose'] = pose.clone()
    return inp


def gt_rgb(n_rays, seed=2):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, n_rays, 3, generator=g)


PERTURB_W, PERTURB_B = 0.004, 0.002   # small enough to keep the geometric-init sphere a surface


def perturb_(model, seed=7, w_std=PERTURB_W, b_std=PERTURB_B, beta=None):
    """"Trained-like" variant of a freshly initialised model (in place, deterministic).

    Geometric init leaves the positional-encoding columns of lin0 and the skip columns of lin4 at zero
    (volsdf/model/network.py:52-58 of the reference), which would hide errors in the PE Jacobian, so
    parity tests also run on weights with seeded Gaussian noise added; `beta` optionally overrides
    density.beta (small beta makes the eval sampler run all 5 iterations).
    """
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters(), key=lambda kv: kv[0]):
            if name.endswith('weight_v') or name.endswith('.weight'):
                p.add_(torch.randn(p.shape, generator=g).to(p.device) * w_std)
            elif name.endswith('bias'):
                p.add_(torch.randn(p.shape, generator=g).to(p.device) * b_std)
            elif name == 'density.beta' and beta is not None:
                p.fill_(beta)
    return model


def mvs_views(n_views=3, dz=48, h=36, w=48, img_res=(72, 96), seed=5):
    """Synthetic stand-in for the MVS stage-0 outputs VolOpt keeps per source view (volsdf/vsdf.py:364-379): a
    probability volume (Dz, H, W), per-pixel depth hypotheses z_mvs (Dz, H, W, increasing along Dz), the view's
    intrinsics at image resolution `img_res` = (height, width) and its camera-to-world pose.  Cameras sit on a circle
    of radius 2.5 around the origin and look at it (the geometric-init sphere), a few degrees apart."""
    g = torch.Generator().manual_seed(seed)
    H, W = img_res
    views = []
    for i in range(n_views):
        ang = (i - (n_views - 1) / 2.0) * 0.25
        c = torch.tensor([2.5 * np.sin(ang), 0.15 * i, -2.5 * np.cos(ang)], dtype=torch.float32)
        fwd = -c / c.norm()
        right = torch.linalg.cross(torch.tensor([0.0, 1.0, 0.0]), fwd)
        right = right / right.norm()
        up = torch.linalg.cross(fwd, right)
        c2w = torch.eye(4, dtype=torch.float32)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, fwd, c
        K = torch.eye(4, dtype=torch.float32)
        K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1] = 1.1 * W, 1.1 * W, W / 2.0 - 0.5, H / 2.0 - 0.5, 0.3
        near = 1.4 + 0.2 * torch.rand(h, w, generator=g)
        far = 3.4 + 0.3 * torch.rand(h, w, generator=g)
        t = torch.linspace(0, 1, dz).view(dz, 1, 1)
        z_mvs = near[None] * (1 - t) + far[None] * t
        cost = torch.softmax(4.0 * torch.randn(dz, h, w, generator=g), 0)
        views.append({'cost': cost.contiguous(), 'z_mvs': z_mvs.contiguous(), 'K': K, 'c2w': c2w})
    return views


def mvs_points(n_rays=64, n_samples=20, seed=6):
    """Ray samples (N, D, 3) around the origin; a fraction falls outside every source frustum / depth range."""
    g = torch.Generator().manual_seed(seed)
    o = torch.tensor([0.0, 0.0, -2.5])
    d = torch.randn(n_rays, 3, generator=g) * torch.tensor([0.35, 0.3, 0.0]) + torch.tensor([0.0, 0.0, 1.0])
    d = d / d.norm(dim=1, keepdim=True)
    z = torch.sort(torch.rand(n_rays, n_samples, generator=g) * 5.5 + 0.05, -1)[0]
    return (o.view(1, 1, 3) + z.unsqueeze(-1) * d.unsqueeze(1)).contiguous()
