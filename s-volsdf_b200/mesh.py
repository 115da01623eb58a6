"""Bulk SDF evaluation on a regular grid for mesh extraction.

The reference evaluates `model.implicit_network(x)[:, 0]` on 100^3 ... 512^3 grids in host-side splits of 100 000
points, moving every split to the CPU (`utils/plots.py:69-76,108-116,188-253`, `eval_vsdf.py:111-150`), and then runs
skimage's marching cubes.  Here the grid goes through the no-grad SDF chain in multi-million-point launches and stays
on the device; marching cubes itself is the caller's (CPU) business, as in the reference.
"""
import torch


@torch.no_grad()
def sdf_grid(model, resolution=100, bound=1.0, chunk=1 << 21, device=None):
    """-> (resolution, resolution, resolution) fp32 tensor of sdf values on [-bound, bound]^3 (x, y, z index order),
    and the 1-D coordinate vector.  `bound` may be a scalar or (xmin, xmax, ymin, ymax, zmin, zmax)."""
    net = model.implicit_network
    dev = device or next(net.parameters()).device
    if isinstance(bound, (int, float)):
        bound = (-bound, bound) * 3
    axes = [torch.linspace(bound[2 * i], bound[2 * i + 1], resolution, device=dev) for i in range(3)]
    out = torch.empty(resolution ** 3, dtype=torch.float32, device=dev)
    n = resolution ** 3
    yz = resolution * resolution
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = torch.arange(lo, hi, device=dev)
        pts = torch.stack([axes[0][idx // yz], axes[1][(idx // resolution) % resolution], axes[2][idx % resolution]], 1)
        out[lo:hi] = net.sdf_only(pts)[:, 0]
    return out.reshape(resolution, resolution, resolution), axes


@torch.no_grad()
def sdf_points(model, points, chunk=1 << 21):
    """sdf (N,) of arbitrary points (N,3) — `plots.get_surface_high_res_mesh`'s refinement queries."""
    net = model.implicit_network
    out = torch.empty(points.shape[0], dtype=torch.float32, device=points.device)
    for lo in range(0, points.shape[0], chunk):
        out[lo:lo + chunk] = net.sdf_only(points[lo:lo + chunk].contiguous())[:, 0]
    return out
