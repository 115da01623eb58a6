"""Times the sampler-sdf forward chain (131072 points, no saves) and the main forward (saves) — used with the
SVS_LIB_PATH variants built by tools/f3_exp.sh to attribute the cycles of tc_fwd3_kernel (GPU box)."""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model
import svolsdf_b200._lib as L
m = build_model('dtu', perturb=True, beta=0.05, device='cuda').set_engine(L.ENGINE_TC_SPLIT).train()
g = torch.Generator().manual_seed(0)
x = (torch.rand(131072, 3, generator=g) * 2 - 1).cuda()
xm = (torch.rand(102400, 3, generator=g) * 2 - 1).cuda()
def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    t1 = timed(lambda: m.implicit_network.get_sdf_vals(x))
t2 = timed(lambda: m.implicit_network.outputs_fused(xm, clamp=100352))
print('%s sampler_sdf %.4f ms  outputs_fwd(+rev) %.4f ms' % (os.environ.get('SVS_LIB_PATH', 'default')[-16:], t1, t2))
