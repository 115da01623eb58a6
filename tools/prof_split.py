"""Short driver for ncu captures: the forward chains of one 1024-ray step on a chosen engine (GPU box)."""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model
import svolsdf_b200._lib as L
eng = {'tc': L.ENGINE_TC, 'split': L.ENGINE_TC_SPLIT}[sys.argv[1] if len(sys.argv) > 1 else 'split']
m = build_model('dtu', perturb=True, beta=0.05, device='cuda').set_engine(eng).train()
g = torch.Generator().manual_seed(0)
x = (torch.rand(131072, 3, generator=g) * 2 - 1).cuda()
xm = (torch.rand(102400, 3, generator=g) * 2 - 1).cuda()
view = torch.nn.functional.normalize(torch.randn(100352, 3, generator=g), dim=1).cuda()
for _ in range(3):
    with torch.no_grad():
        m.implicit_network.get_sdf_vals(x)
    y, s, gq = m.implicit_network.outputs_fused(xm, clamp=100352)
    r = m.rendering_network(xm[:100352], gq[:100352], view, y[:100352], _feat_col=1)
torch.cuda.synchronize()
