"""Full-image rendering (`VolOpt.render_step`, volsdf/vsdf.py:237-262; `eval_vsdf.evaluate`, eval_vsdf.py:216-228).

The reference splits the image into chunks of 500/512 rays with `utils.split_input`, calls the model once per chunk
and moves six tensors to the CPU after every call (885 model calls for 576x768).  Here the frame is ray-sharded over
the ranks, each rank walks its rays in large chunks (thousands of 128-point tiles per launch) and the three image
planes stay on the device; one all-gather (28 B/ray) assembles the frame.

`chunk` is also the sampler's convergence group (ray_sampler.py:136 decides per model call whether EVERY ray of the
call gets another 128 samples): `chunk=512` reproduces eval_vsdf.py's grouping exactly, larger chunks only add
samples to rays the reference would have stopped early.
"""
import torch
import torch.distributed as dist

from . import dist as sdist


@torch.no_grad()
def render_rays(model, intrinsics, pose, uv, chunk=16384, extra=None):
    """uv (1,R,2) on the device -> dict of rgb_values (R,3), depth_values (R,1), normal_map (R,3)"""
    model.eval()
    R = uv.shape[1]
    out = {'rgb_values': [], 'depth_values': [], 'normal_map': []}
    iters = []
    for lo in range(0, R, chunk):
        inp = {'intrinsics': intrinsics, 'pose': pose, 'uv': uv[:, lo:lo + chunk].contiguous()}
        if extra:
            inp.update(extra)
        o = model(inp)
        for k in out:
            out[k].append(o[k])
        iters.append(model.ray_sampler.last_iters)
    res = {k: torch.cat(v, 0) for k, v in out.items()}
    res['sampler_iters'] = iters
    return res


@torch.no_grad()
def render_image(model, intrinsics, pose, uv, chunk=16384, rank=0, world=1, extra=None):
    """Renders the rays `uv` (1,R,2) ray-sharded over `world` ranks; every rank returns the full planes."""
    R = uv.shape[1]
    lo, hi = sdist.shard_range(R, rank, world)
    part = render_rays(model, intrinsics, pose, uv[:, lo:hi], chunk=chunk, extra=extra)
    if world == 1:
        return part
    packed = torch.cat([part['rgb_values'], part['depth_values'], part['normal_map']], 1)   # (r, 7)
    sizes = [sdist.shard_range(R, r, world) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    buf = torch.zeros(nmax, 7, device=packed.device)
    buf[:packed.shape[0]] = packed
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    full = torch.cat([g[:b - a] for g, (a, b) in zip(gathered, sizes)], 0)
    return {'rgb_values': full[:, :3], 'depth_values': full[:, 3:4], 'normal_map': full[:, 4:7],
            'sampler_iters': part['sampler_iters']}
