"""Where does the error of d loss / d beta come from?  Compositor backward per 32-ray chunk vs fp64 autograd (GPU box)."""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model
from oracle import volsdf_oracle as O
import svolsdf_b200._lib as L
import svolsdf_b200.scene as S
from svolsdf_b200 import functional as F

DEV = 'cuda'
R = 1024
beta = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
m = build_model('dtu', perturb=True, beta=beta, device=DEV).train()
inp = {k: v.to(DEV) for k, v in S.make_input('dtu', R).items()}
gt = S.gt_rgb(R).reshape(-1, 3).to(DEV)
torch.manual_seed(321)
out = m(inp, fast=1)
z, _ = m.last_z
# recompute sdf / rgb for the main samples (same kernels)
ray_dirs, cam_loc, depth_scale = m._rays(inp['uv'], inp['pose'], inp['intrinsics'])
pts = F.ray_points(cam_loc, ray_dirs, z).reshape(-1, 3)
with torch.no_grad():
    y, sdf, g = m.implicit_network.outputs_fused(pts, clamp=True)
    rgb = m.rendering_network(pts, g, ray_dirs.unsqueeze(1).expand(R, 98, 3).reshape(-1, 3), y, _feat_col=1)
sdf = sdf.reshape(R, 98)
rgb = rgb.reshape(R, 98, 3)
tot_gpu = {0: 0.0, L.COMP_FAST: 0.0}
tot_ref = 0.0
worst = []
for lo in range(0, R, 32):
    sl = slice(lo, lo + 32)
    # fp64 reference
    bp = m.density.beta.detach().double().cpu().clone().requires_grad_(True)
    s64 = sdf[sl].double().cpu().clone().requires_grad_(True)
    b = O.get_beta(bp, 1e-4)
    w = O.volume_rendering(z[sl].double().cpu(), s64.reshape(-1, 1), b)
    rv, dv, _ = O.composite(w, rgb[sl].double().cpu(), z[sl].double().cpu(), depth_scale[sl].double().cpu())
    ((rv - gt[sl].double().cpu()).abs().sum() / (3 * R)).backward()
    ref = float(bp.grad)
    tot_ref += ref
    row = [lo, ref]
    for flags in (0, L.COMP_FAST):
        bpg = m.density.beta.detach().clone().requires_grad_(True)
        sg = sdf[sl].clone().requires_grad_(True)
        wts, rvals, dvals, _, _ = F.composite(z[sl].contiguous(), sg, rgb[sl].contiguous(), bpg, 1e-4, depth_scale[sl].contiguous(), flags=flags)
        ((rvals - gt[sl]).abs().sum() / (3 * R)).backward()
        got = float(bpg.grad)
        tot_gpu[flags] += got
        row += [got, float((sg.grad.double().cpu() - s64.grad).abs().max() / s64.grad.abs().max())]
    worst.append(row)
worst.sort(key=lambda r: -abs(r[2] - r[1]))
print('total ref %.6e  canonical %.6e  fast %.6e' % (tot_ref, tot_gpu[0], tot_gpu[L.COMP_FAST]))
for r in worst[:6]:
    print('chunk %4d ref %.5e canonical %.5e (d_sdf rel %.2e) fast %.5e (d_sdf rel %.2e)' % tuple(r))
