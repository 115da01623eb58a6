"""Per-kernel device time of a partial 1600x1200 frame render (the first N rays of the grid, beta = 0.01, 512-ray groups):
which kernels the frame's time goes to, and wall time vs summed kernel time (host bubbles).  GPU box, measurement only."""
import os, sys, time, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import svolsdf_b200._lib as L
import svolsdf_b200.conf as C
import svolsdf_b200.scene as S
from svolsdf_b200.model.network import VolSDFNetwork
from svolsdf_b200.render import render_image
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
torch.manual_seed(0)
model = VolSDFNetwork(C.dtu_model_conf())
S.perturb_(model, w_std=0.0, b_std=0.0, beta=0.01)
model = model.cuda().eval().set_engine(L.ENGINE_TC_SPLIT)
inp = S.make_input('dtu', 1600 * 1200, width=1600, height=1200, pixels='grid')
K_, pose, uv = inp['intrinsics'].cuda(), inp['pose'].cuda(), inp['uv'][:, 600 * 1600:600 * 1600 + N].cuda()
render_image(model, K_, pose, uv, chunk=16384, group=512)
torch.cuda.synchronize()
t0 = time.time()
render_image(model, K_, pose, uv, chunk=16384, group=512)
torch.cuda.synchronize()
wall = (time.time() - t0) * 1e3
L.prof_enable(True)
out = render_image(model, K_, pose, uv, chunk=16384, group=512)
torch.cuda.synchronize()
prof = L.prof_collect()
L.prof_enable(False)
tot = sum(v['ms'] for v in prof.values())
print('rays %d  wall %.1f ms  library kernels %.1f ms  sampler iterations (mean) %.2f' % (N, wall, tot, sum(out['sampler_iters']) / len(out['sampler_iters'])))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    print('  %-22s %8.2f ms  %5.1f %%  x%d' % (k, v['ms'], 100 * v['ms'] / tot, v['launches']))
