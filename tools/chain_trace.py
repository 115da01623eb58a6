"""Reads the clock64 trace of CTA 0 (third tile) of one tc_chain_kernel instantiation (library built with
-DSVS_CHAIN_TRACE=<prologue id>: 4 tangent, 5 backward, 1 reverse, 3 rendering-net backward; tools/f3_exp.sh) during a
1024-ray train step and prints the timeline of the MMA issuer, aux producer, store lane and three epilogue warps
(GPU box, measurement only)."""
import os, sys, ctypes as C, warnings
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
lib = C.CDLL(L.LIB_PATH)
buf = (C.c_longlong * 16384)()
def step():
    m.zero_grad()
    out = m(inp, fast=1)
    loss = (out['rgb_values'] - gt).abs().mean() + 0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
    loss.backward()
for _ in range(3): step()
lib.svs_dbg_f3_trace(buf, 16384)
step()
n = lib.svs_dbg_f3_trace(buf, 16384)
ev = []
for i in range(n):
    v = buf[i]
    if v == 0: continue
    ev.append((v & 0xFFFFFFFFFF, i // 1024, (v >> 56) & 255, (v >> 48) & 255, (v >> 40) & 255))
ev.sort()
t0 = ev[0][0]
names = {1: 'mma step start', 2: 'mma w_full', 3: 'mma a_ready', 6: 'mma commit acc', 7: 'epi acc_full', 8: 'epi piece start', 9: 'epi tmem ok', 13: 'epi aux ready', 10: 'epi computed',
         14: 'epi s_free ok', 12: 'epi arrived', 20: 'aux slot free', 40: 'tile top', 41: 'row loads done', 42: 'tile barrier', 43: 'A free / loaded', 44: 'dy rows stored', 45: 'dy col0 done', 46: 'pro piece arrived', 47: 'pegrad pieces done', 48: 'pegrad bar1', 49: 'pegrad computed', 50: 'pegrad bar2', 30: 'store block ready', 31: 'store read done'}
who = {0: 'MMA', 1: 'E0', 2: 'E5', 3: 'E15', 4: 'AUX', 5: 'ST'}
print('events', len(ev))
for t, ln, tag, a, b in ev:
    print('%8d  %-4s %-18s s=%d  %d' % (t - t0, who.get(ln, ln), names.get(tag, tag), a, b))
