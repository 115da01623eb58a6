"""Reads the clock64 trace of CTA 0 of tc_fwd3_kernel (library built with -DSVS_F3_TRACE, tools/f3_exp.sh) and prints
the per-layer timeline of the MMA issuer and of two epilogue warps (GPU box, measurement only)."""
import os, sys, ctypes as C, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
warnings.filterwarnings('ignore')
from helpers import build_model
import svolsdf_b200._lib as L
m = build_model('dtu', perturb=True, beta=0.05, device='cuda').set_engine(L.ENGINE_TC_SPLIT).train()
g = torch.Generator().manual_seed(0)
x = (torch.rand(131072, 3, generator=g) * 2 - 1).cuda()
lib = C.CDLL(L.LIB_PATH)
buf = (C.c_longlong * 16384)()
with torch.no_grad():
    for _ in range(3): m.implicit_network.get_sdf_vals(x)
    lib.svs_dbg_f3_trace(buf, 16384)
    m.implicit_network.get_sdf_vals(x)
n = lib.svs_dbg_f3_trace(buf, 16384)
ev = []
for i in range(n):
    v = buf[i]
    if v == 0: continue
    ev.append((v & 0xFFFFFFFFFF, (v >> 56) & 255, (v >> 48) & 255, (v >> 40) & 255))
ev.sort()
t0 = ev[0][0]
names = {1: 'mma layer start', 2: 'mma w_full lo', 3: 'mma a_ready', 4: 'mma w_full hi', 5: 'mma w_full main', 6: 'mma commit acc', 7: 'epi acc_full', 8: 'epi piece done', 9: 'epi tmem ld ok', 10: 'epi computed', 11: 'epi stored', 12: 'epi fenced'}
print('events', len(ev))
lim = int(sys.argv[1]) if len(sys.argv) > 1 else 400
only = sys.argv[2] if len(sys.argv) > 2 else ''
for t, tag, a, b in ev[:lim]:
    if only and not names.get(tag, '').startswith(only): continue
    print('%8d  %-16s s=%d  %d' % (t - t0, names.get(tag, tag), a, b))
