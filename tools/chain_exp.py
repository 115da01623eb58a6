"""Per-kernel device times of eager 1024-ray train steps for the library in SVS_LIB_PATH (A/B runs of build variants:
tools/f3_exp.sh).  GPU box, measurement only."""
import os, sys, warnings, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model
import svolsdf_b200._lib as L
import svolsdf_b200.scene as S
R = 1024
m = build_model('dtu', perturb=True, beta=0.05, device='cuda').set_engine(L.ENGINE_TC_SPLIT).train()
inp = {k: v.cuda() for k, v in S.make_input('dtu', R).items()}
gt = S.gt_rgb(R).reshape(-1, 3).cuda()
def step():
    m.zero_grad()
    out = m(inp, fast=1)
    loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
L.prof_enable(True)
for _ in range(5): step()
torch.cuda.synchronize()
prof = L.prof_collect()
L.prof_enable(False)
print(os.environ.get('SVS_LIB_PATH', 'default')[-14:], ' '.join('%s %.3f' % (k.replace('mlp_tc_', ''), v['ms'] / 5) for k, v in sorted(prof.items()) if 'mlp_tc' in k))
