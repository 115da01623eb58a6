"""Full-image rendering (`VolOpt.render_step`, volsdf/vsdf.py:237-262; `eval_vsdf.evaluate`, eval_vsdf.py:216-228).

The reference splits the image into chunks of 500/512 rays with `utils.split_input`, calls the model once per chunk
and moves six tensors to the CPU after every call (885 model calls for 576x768).  Here the frame is ray-sharded over
the ranks, each rank walks its rays in large chunks (thousands of 128-point tiles per launch) and the three image
planes stay on the device; one all-gather (28 B/ray) assembles the frame.

`group` is the sampler's convergence group: ray_sampler.py:136 decides per MODEL CALL whether every ray of the call
gets another 128 samples, and the reference's callers feed the model `split_n_pixels` rays per call (500 in
vsdf.py:246-262, 512 in eval_vsdf.py:216-228).  A launch chunk here holds chunk / group such groups, each converging on
its own (ErrorBoundSampler.group_size), so the frame equals what the reference's chunked loop computes ray for ray while
every kernel still sees thousands of tiles.
"""
import torch
import torch.distributed as dist

from . import dist as sdist


@torch.no_grad()
def render_rays(model, intrinsics, pose, uv, chunk=16384, extra=None, group=512):
    """uv (1,R,2) on the device -> dict of rgb_values (R,3), depth_values (R,1), normal_map (R,3).  `chunk` rays per
    launch (a multiple of `group`), `group` rays per convergence group (None: one group per launch chunk)."""
    model.eval()
    R = uv.shape[1]
    if group:
        chunk = max(group, chunk // group * group)
    out = {'rgb_values': [], 'depth_values': [], 'normal_map': []}
    iters = []
    old = model.ray_sampler.group_size
    model.ray_sampler.group_size = group
    try:
        for lo in range(0, R, chunk):
            inp = {'intrinsics': intrinsics, 'pose': pose, 'uv': uv[:, lo:lo + chunk].contiguous()}
            if extra:
                inp.update(extra)
            o = model(inp)
            for k in out:
                out[k].append(o[k])
            gi = model.ray_sampler.last_group_iters
            iters.extend(gi if gi is not None else [model.ray_sampler.last_iters])
    finally:
        model.ray_sampler.group_size = old
    res = {k: torch.cat(v, 0) for k, v in out.items()}
    res['sampler_iters'] = iters
    return res


@torch.no_grad()
def render_image(model, intrinsics, pose, uv, chunk=16384, rank=0, world=1, extra=None, group=512):
    """Renders the rays `uv` (1,R,2) ray-sharded over `world` ranks; every rank returns the full planes."""
    R = uv.shape[1]
    lo, hi = sdist.shard_range(R, rank, world, align=group or 1)
    part = render_rays(model, intrinsics, pose, uv[:, lo:hi], chunk=chunk, extra=extra, group=group)
    if world == 1:
        return part
    packed = torch.cat([part['rgb_values'], part['depth_values'], part['normal_map']], 1)   # (r, 7)
    sizes = [sdist.shard_range(R, r, world, align=group or 1) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    buf = torch.zeros(nmax, 7, device=packed.device)
    buf[:packed.shape[0]] = packed
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    full = torch.cat([g[:b - a] for g, (a, b) in zip(gathered, sizes)], 0)
    return {'rgb_values': full[:, :3], 'depth_values': full[:, 3:4], 'normal_map': full[:, 4:7],
            'sampler_iters': part['sampler_iters']}
