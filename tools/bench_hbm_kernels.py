"""HBM roofline of the sampler and compositor kernels at a ray count where they are bandwidth- rather than
launch-bound (R >= 65536).  Algorithmic bytes per ray are SURVEY.md §8(d)'s figures.
    python tools/bench_hbm_kernels.py [R]   -> one JSON line per kernel (+ writes gpurun_out/hbm_kernels.json)"""
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import svolsdf_b200._lib as L  # noqa: E402
from svolsdf_b200 import functional as F  # noqa: E402
from svolsdf_b200.model.ray_sampler import _cfg, _linspace  # noqa: E402

dev = 'cuda'
R = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
S = 98
peak = 6459.0
p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p))['hbm_gbs']


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator().manual_seed(0)
z = torch.sort(torch.rand(R, S, generator=g) * 5.0 + 0.5, dim=1)[0].to(dev)
sdf = ((torch.rand(R, S, generator=g) - 0.3) * 0.5).to(dev)
rgb = torch.rand(R, S, 3, generator=g).to(dev)
beta = torch.tensor([0.05], device=dev)
ds = torch.ones(R, 1, device=dev)
res = []

# compositor forward (training layout: no normals): reads z, sdf, rgb; writes weights, rgb_values, depth_values
w = torch.empty(R, S, device=dev); rv = torch.empty(R, 3, device=dev); dv = torch.empty(R, 1, device=dev)
for fl, tag in ((0, 'exact'), (L.COMP_FAST, 'fast')):
    ms = timeit(lambda: L.call('svs_composite_forward', z.data_ptr(), sdf.data_ptr(), rgb.data_ptr(), None, beta.data_ptr(), 1e-4,
                               ds.data_ptr(), None, R, S, fl, w.data_ptr(), rv.data_ptr(), dv.data_ptr(), None, None, L.stream()))
    res.append(('composite_fwd_' + tag, 2368.0, ms))
# compositor backward: reads dW, d(rgb, depth), forward inputs; writes d_sdf, d_rgb
dwt = torch.rand(R, S, device=dev); drv = torch.rand(R, 3, device=dev); ddv = torch.rand(R, 1, device=dev)
d_sdf = torch.empty(R, S, device=dev); d_rgb = torch.empty(R, S, 3, device=dev); d_beta = torch.zeros(1, device=dev)
for fl, tag in ((0, 'exact'), (L.COMP_FAST, 'fast')):
    ms = timeit(lambda: L.call('svs_composite_backward', z.data_ptr(), sdf.data_ptr(), rgb.data_ptr(), beta.data_ptr(), 1e-4,
                               ds.data_ptr(), None, R, S, fl, drv.data_ptr(), ddv.data_ptr(), dwt.data_ptr(), None,
                               d_sdf.data_ptr(), d_rgb.data_ptr(), d_beta.data_ptr(), L.stream()))
    res.append(('composite_bwd_' + tag, 3936.0, ms))

# sampler, training iteration: init (t_rand 512 B in, z 512 B out), bound (z, sdf in; beta out), resample (u in, samples out),
# finalize (z_98 out): 2212 B/ray in total
for exact in (1, 0):
    cfg = _cfg(0.0, 6.0, 0.1, 0.0, 10, bool(exact))
    n = 128
    t_rand = torch.rand(R, n, generator=g).to(dev)
    zz = torch.empty(R, n, device=dev); bb = torch.empty(R, device=dev)
    # sdf rows of the synthetic DTU scene: unit sphere of radius 0.6 seen from (0, 0, -2.5) through the DTU
    # pinhole, with the bounding-sphere clamp min(sdf, 20 (3 - |x|)) of the benchmark model; samples = sampler_init's z
    import svolsdf_b200.scene as SC
    from svolsdf_b200 import functional as FF
    inp = SC.make_input('dtu', R, pixels='perm')
    dirs, cam, _ = FF.raygen(inp['uv'][0].to(dev), inp['pose'][0].to(dev), inp['intrinsics'][0].to(dev))
    L.call('svs_sampler_init', cfg, R, n, lin.data_ptr() if False else _linspace(n, dev).data_ptr(), t_rand.data_ptr(), None, zz.data_ptr(), bb.data_ptr(), L.stream())
    pts = cam[:, None, :] + zz[:, :, None] * dirs[:, None, :]
    nr = pts.norm(dim=-1)
    sdf_new = torch.minimum(nr - 0.6, 20.0 * (3.0 - nr)).contiguous()
    del pts, nr
    sdf_m = torch.empty(R, n, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
    u = torch.rand(R, 64, generator=g).to(dev); samples = torch.empty(R, 64, device=dev)
    extra = torch.randperm(n)[:32].to(torch.int32).to(dev); eik = torch.randint(98, (R,)).to(dev)
    zf = torch.empty(R, 98, device=dev); ze = torch.empty(R, 1, device=dev)
    lin = _linspace(n, dev)

    def it():
        st = L.stream()
        L.call('svs_sampler_init', cfg, R, n, lin.data_ptr(), t_rand.data_ptr(), None, zz.data_ptr(), bb.data_ptr(), st)
        L.call('svs_sampler_bound', cfg, R, n, n, zz.data_ptr(), None, sdf_new.data_ptr(), None, sdf_m.data_ptr(),
               beta.data_ptr(), 1e-4, bb.data_ptr(), flag.data_ptr(), st)
        L.call('svs_sampler_resample', cfg, R, n, 64, 0, zz.data_ptr(), sdf_m.data_ptr(), bb.data_ptr(), u.data_ptr(), 1,
               samples.data_ptr(), None, None, None, st)
        L.call('svs_sampler_finalize', cfg, R, n, 64, zz.data_ptr(), samples.data_ptr(), extra.data_ptr(), 32, None,
               eik.data_ptr(), zf.data_ptr(), ze.data_ptr(), st)
    ms = timeit(it)
    res.append(('sampler_train_iteration_%s' % ('exact_fp64' if exact else 'fast_fp32'), 2212.0, ms))
    L.prof_enable(True)
    for _ in range(5):
        it()
    torch.cuda.synchronize()
    pr = L.prof_collect()
    L.prof_enable(False)
    print(json.dumps({'sampler_breakdown_ms_per_launch': {k: v['ms'] / v['launches'] for k, v in pr.items()}, 'exact': exact}))

# MVS cost lookup (VolOpt.cost_mapping): 3 source views of 48 x 288 x 384 (the paper configuration, vsdf.py:369), 98
# samples per ray.  Bytes per ray: 98 samples x (12 in + 9 out + 3 views x 16 taps x 4 B gathered from L2-resident volumes)
if R <= 262144:
    import svolsdf_b200.scene as S
    from svolsdf_b200.mvs import CostMapper
    Rm = min(R, 65536)
    views = S.mvs_views(n_views=3, dz=48, h=288, w=384, img_res=(1152, 1536), seed=9)
    cm = CostMapper([v['cost'][None] for v in views], [v['z_mvs'][None] for v in views], [v['K'] for v in views],
                    [v['c2w'] for v in views], [25, 22, 28], (1152, 1536))
    xyz = S.mvs_points(Rm, 98, seed=10).to(dev)
    zt = torch.zeros(Rm, 98, device=dev)
    own = torch.tensor([22])
    ms = timeit(lambda: cm(zt, own, xyz)) * (R / Rm)
    res.append(('cost_mapping_3views', 98 * (21.0 + 3 * 64.0), ms))

out = []
for name, bpr, ms in res:
    gbs = bpr * R / (ms * 1e-3) / 1e9
    row = {'kernel': name, 'rays': R, 'algorithmic_bytes_per_ray': bpr, 'ms': ms, 'achieved_gbs': gbs, 'peak_gbs': peak,
           'frac': gbs / peak}
    out.append(row)
    print(json.dumps(row))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'hbm_kernels.json'), 'w'), indent=1)
