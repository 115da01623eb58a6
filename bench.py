#!/usr/bin/env python
"""bench.py — headline benchmark of the VolSDF hot path (BASELINE.json: rays/sec of the fwd+bwd train step).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload (`config.workload`): configs[1] of BASELINE.json — DTU-shaped VolSDF training step (sampler 128 +
main pass 98 + eikonal 2 SDF evaluations per ray, L1 + 0.1*eikonal loss, backward incl. double backward,
grad clip + Adam as in volsdf/vsdf.py:196-219), 1024 rays per GPU, random-init networks, synthetic camera.
One "step" = one such optimisation step.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

import torch  # noqa: E402

FLOP_PER_RAY_TRAIN = 920586240.0   # 128F + 98(6F+3F_r) + 12F, SURVEY.md §8d / Appendix B


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--rays', type=int, default=1024, help='rays per GPU per step')
    ap.add_argument('--engine', default='auto', choices=['auto', 'fp32', 'tc', 'tc_split'],
                    help='auto = tc_split: the tcgen05 engine that meets the parity contract')
    ap.add_argument('--cpu-rays', type=int, default=1024, help='rays per step of the bounded CPU-baseline sample (default: the headline config)')
    ap.add_argument('--cpu-steps', type=int, default=6, help='timed steps of the CPU-baseline sample (1 warm-up before them; ~1.8 s each at 1024 rays)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='all', choices=['all', 'train', 'frame'],
                    help="all (default): the headline train step (BASELINE configs[1]) plus `frame` (configs[2]) and `large_batch` "
                         "(configs[4]) legs in the same JSON line; train / frame: only that one")
    ap.add_argument('--group', type=int, default=512, help='frame: rays per sampler convergence group (eval_vsdf.py: 512)')
    ap.add_argument('--large-rays', type=int, default=8192, help='large_batch leg: rays per GPU (8 GPUs: 65536 per step)')
    ap.add_argument('--ref-rays', type=int, default=None, help='--impl reference: rays per step (default: --rays, the same config)')
    ap.add_argument('--skip', default='', help='comma list of legs to skip: frame,large,parity,gpu_eager,engines,mvs')
    ap.add_argument('--chunk', type=int, default=16384, help='frame workload: rays per model call (= sampler convergence group)')
    ap.add_argument('--beta', type=float, default=None, help='frame workload: density.beta override (0.01 = trained-like, 5 sampler iterations)')
    ap.add_argument('--scene', default='dtu', choices=['dtu', 'bmvs'],
                    help='train workload: dtu = BASELINE configs[1] (the headline); bmvs = configs[3], VolSDFNetworkBG with the '
                         'inverted-sphere background networks at 768x576')
    ap.add_argument('--rng', default='auto', choices=['auto', 'global', 'local'],
                    help='multi-GPU host random draws: global = every rank draws the whole batch and keeps its rows (sharded run '
                         'bit-identical per ray to the unsharded one, host work grows with N); local = per-rank streams (data-parallel '
                         'semantics, host work constant); auto = local when N > 1')
    ap.add_argument('--eager', action='store_true', help='issue every step from Python instead of replaying a CUDA graph')
    ap.add_argument('--allreduce', default='peer', choices=['peer', 'nccl'],
                    help='N > 1: peer = one kernel over NVLink peer memory (all-reduce + clip + Adam); nccl = NCCL all-reduce + fused step')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops_burst': d['bf16_tflops'],
                'tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        """Launches nvidia-smi (10 ms period).  Call it BEFORE the warm-up: the tool needs ~0.1 s to deliver its first
        line, longer than a 20-step timed region; mark()/stop() then keep the samples that fall inside the region."""
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '10'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark(self):
        """Call at the start and at the end of the timed region."""
        if self.t0 is None:
            self.t0 = time.monotonic()
        else:
            self.t1 = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        window = 'timed region'
        t0, t1 = self.t0, self.t1 if self.t1 is not None else time.monotonic()
        lines = [ln for ts, ln in self.lines if t0 is None or t0 <= ts <= t1 + 0.015]
        if not lines and t0 is not None:
            # a region shorter than the sampling period: the samples taken under the same load right after it (the
            # end-to-end region replays the same step) stand in
            lines = [ln for ts, ln in self.lines if ts >= t0]
            window = 'timed + end-to-end regions'
        self.lines = [(0.0, ln) for ln in lines]
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for _ts, ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm), 'window': window}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host CPU (the real reference when /root/reference is
# mounted, else the oracle port — the reference is pure Python/PyTorch and cannot travel to the GPU box)
# ------------------------------------------------------------------------------------------------------

def cpu_train_steps(n_rays, steps, warmup, threads=None):
    import svolsdf_b200.conf as C
    import svolsdf_b200.scene as S
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 for every rank)
    torch.set_num_threads(threads or max(1, min(os.cpu_count() or 1, 64)))
    kind = 'port'
    inp = S.make_input('dtu', n_rays)
    gt = S.gt_rgb(n_rays)
    step = None
    try:
        from oracle import ref_import
        if ref_import.available():
            ns = ref_import.load()
            torch.manual_seed(0)
            model = ns.network.VolSDFNetwork(C.dtu_model_conf()).train()
            opt = torch.optim.Adam(model.parameters(), lr=5e-4)
            kind = 'reference'

            def step():
                out = model(inp, fast=1)
                loss = (out['rgb_values'] - gt.reshape(-1, 3)).abs().mean() + \
                    0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()
                opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
                opt.step()
                return float(loss)
    except Exception:
        step = None
    if step is None:
        from oracle import volsdf_oracle as O
        from svolsdf_b200.model.network import VolSDFNetwork
        torch.manual_seed(0)
        m = VolSDFNetwork(C.dtu_model_conf())
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        opt = torch.optim.Adam(list(sd.values()), lr=5e-4)
        conf = C.dtu_model_conf()

        def step():
            out = O.volsdf_forward(sd, conf, inp, True, fast=1)
            loss = O.volsdf_loss(out, gt)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(sd.values()), 1.0)
            opt.step()
            return float(loss)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {'value': n_rays / dt, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': kind,
            'sample': '%d-ray DTU train step (fwd+loss+bwd+clip+Adam), %d warm-up + %d timed, %.2f s/step'
                      % (n_rays, warmup, steps, dt), 'ms_per_step': dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = args.ref_rays or args.rays          # the same config as this repo's arm: 1024 rays per step
    steps = max(1, args.steps)
    r = cpu_train_steps(n, steps, max(0, args.warmup))
    line = {
        'metric': 'rays/sec (fwd+bwd train step)', 'value': r['value'], 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': max(0, args.warmup), 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': 'DTU VolSDF train step fwd+bwd+eikonal (BASELINE configs[1]): sampler 128 + main 98 + '
                               'eikonal 2 SDF evals/ray, L1+0.1*eik loss, clip+Adam — the reference algorithm on the host CPU cores',
                   'rays_per_gpu': n, 'global_rays': n, 'rays_per_step': n},
        'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': r['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------

ENGINE_TEXT = {
    0: ('f32', 'fp32 SIMT (parity mode)'),
    1: ('f16', 'tcgen05 kind::f16, single fp16 operands, fp32 TMEM accumulate (fastest; sdf ~1e-3: OUTSIDE the 1e-3 depth contract)'),
    2: ('f16', 'tcgen05 kind::f16, split fp16 hi+lo operands (3 MMAs/layer) in every forward chain, single fp16 in the backward '
               'chains, fp32 TMEM accumulate: rgb/depth <= 1e-3, parameter gradients <= 1e-2 vs the fp64 oracle (tests/test_gpu_split.py)'),
}


def pick_engine(args, L):
    return {'auto': L.ENGINE_TC_SPLIT, 'tc_split': L.ENGINE_TC_SPLIT, 'tc': L.ENGINE_TC, 'fp32': L.ENGINE_FP32}[args.engine]


class TrainRig(object):
    """One DTU / BlendedMVS train step (forward + loss + backward [+ gradient all-reduce] + clip + Adam) at R rays per GPU:
    eager closure, CUDA-graph replay, and the end-to-end variant with host inputs / host RNG / loss read-back."""

    def __init__(self, args, R, engine, world, rank, dev, n_tapes, scene):
        import svolsdf_b200._lib as L
        import svolsdf_b200.conf as C
        import svolsdf_b200.scene as S
        from svolsdf_b200 import dist as sdist
        from svolsdf_b200.model.network import VolSDFNetwork
        from svolsdf_b200.model.ray_sampler import RecordedRng, RefRng, TapeRng
        from svolsdf_b200.optim import FusedAdam
        self.L, self.R, self.world, self.dev = L, R, world, dev
        self.Rg = Rg = R * world
        torch.manual_seed(0)
        if scene == 'bmvs':
            from svolsdf_b200.model.network_bg import VolSDFNetworkBG
            model = VolSDFNetworkBG(C.bmvs_model_conf()).to(dev).train().set_engine(engine)
        else:
            model = VolSDFNetwork(C.dtu_model_conf()).to(dev).train().set_engine(engine)
        self.model = model
        # clip_grad_norm_(1.0) + NaN guard + Adam of vsdf.py:214-219 as one fused step (svolsdf_b200.optim.FusedAdam)
        # N > 1: the gradient all-reduce is fused into that step — ONE kernel reads every rank's flat gradient buffer over
        # NVLink peer memory (svs_adam_step_allreduce; --allreduce nccl = round 1's NCCL all-reduce + separate step)
        self.allreduce = 'none'
        peer = None
        if world > 1 and args.allreduce == 'peer':
            try:
                peer = sdist.PeerGradBuffer(model.parameters())
                self.allreduce = 'peer-memory kernel fused with clip+Adam (svs_adam_step_allreduce)'
            except Exception as e:
                peer = None
                self.allreduce = 'nccl (symmetric memory unavailable: %s)' % (str(e).splitlines()[0][:100],)
        elif world > 1:
            self.allreduce = 'nccl all-reduce of the flat fp32 gradients + fused clip+Adam'
        self.opt = opt = FusedAdam(model.parameters(), lr=5e-4, max_grad_norm=1.0, peer=peer)
        self.reducer = reducer = sdist.GradAllReducer(model.parameters()) if (world > 1 and peer is None) else None
        inp_host = S.make_input(scene, Rg, pixels='perm' if Rg > 4096 else 'random')
        gt_host = S.gt_rgb(Rg)
        lo, hi = sdist.shard_range(Rg, rank, world)
        inp_host = sdist.shard_input(inp_host, rank, world)
        gt_host = gt_host[:, lo:hi].contiguous()
        self.inp_host, self.gt_host = inp_host, gt_host
        self.inp_pin = {k: v.pin_memory() for k, v in inp_host.items()}
        self.gt_pin = gt_host.pin_memory()
        self.inp_dev = {k: v.to(dev) for k, v in inp_host.items()}
        self.gt_dev = gt_host.to(dev)
        self.rng_mode = 'single' if world == 1 else ('global' if args.rng == 'global' else 'local')
        rng_mode = self.rng_mode

        def make_rng():
            return sdist.ShardedRng(dev, Rg, lo, hi) if rng_mode == 'global' else RefRng(dev)
        self.make_rng = make_rng
        # random draws of every step, made in the reference's order and uploaded BEFORE the timed region
        torch.manual_seed(1234 + (rank if rng_mode == 'local' else 0))
        self.tapes = []
        for _ in range(n_tapes):
            tr = TapeRng(make_rng())
            model.rng_source = tr
            with torch.no_grad():   # a dry forward only to make the draws in the reference's order (not timed)
                model(self.inp_dev, fast=1)
            self.tapes.append(tr.tape)
        model.rng_source = None
        torch.cuda.synchronize()
        self.RecordedRng = RecordedRng
        l0 = L.launch_count()
        self.step_eager(0)
        torch.cuda.synchronize()
        self.launches_per_step = L.launch_count() - l0
        self.graphed, self.step_mode = None, 'eager'
        if not args.eager:
            try:
                from svolsdf_b200.train import GraphedTrainStep
                self.graphed = GraphedTrainStep(model, opt, self.loss_of, self.inp_dev, self.gt_dev, grad_clip=0.0,
                                                reducer=reducer, world=world, make_rng=make_rng)
                self.step_mode = 'cuda_graph'
            except Exception as e:   # keep the run alive, say so in the JSON line
                self.graphed, self.step_mode = None, 'eager (graph capture failed: %s)' % (str(e).splitlines()[0][:120],)
                model.rng_source = None
        self.pending = None

    @staticmethod
    def loss_of(out, gt):
        return (out['rgb_values'] - gt.reshape(-1, 3)).abs().mean() + \
            0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()

    def finish(self, loss):
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.allreduce_(self.world)
        self.opt.step()      # gradient clipping to norm 1.0 happens inside the fused step

    def step_eager(self, i):
        self.model.rng_source = self.RecordedRng(self.dev, self.tapes[i % len(self.tapes)])
        out = self.model(self.inp_dev, fast=1)
        self.finish(self.loss_of(out, self.gt_dev))

    def step_device(self, i):
        if self.graphed is not None:
            self.graphed(draws=self.tapes[i % len(self.tapes)])     # inputs + this step's random draws already resident in HBM
        else:
            self.step_eager(i)

    def step_e2e(self):
        if self.graphed is not None:
            # host RNG in the reference's order -> pinned -> device.  The draws of step i+1 are made on the CPU while
            # the GPU replays step i (they only depend on the CPU generator), then the loss of step i is read back.
            draws = self.pending if self.pending is not None else self.graphed.draw()
            loss = self.graphed(self.inp_pin, self.gt_pin, draws)
            self.pending = self.graphed.draw()
            return float(loss.item()), self.graphed.h2d_bytes_rng
        self.model.rng_source = self.make_rng()
        inp = {k: v.to(self.dev, non_blocking=True) for k, v in self.inp_pin.items()}
        gt = self.gt_pin.to(self.dev, non_blocking=True)
        out = self.model(inp, fast=1)
        loss = self.loss_of(out, gt)
        self.finish(loss)
        return float(loss.item()), self.model.rng_source.h2d_bytes   # device->host read of the step's result

    def release(self):
        self.graphed = None
        self.model.rng_source = None


def timed(fn, K, barrier, max_over_ranks):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        fn(i)
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / K


def parity_leg(engine, R=1024):
    """Measured errors of the benchmarked engine at the benchmarked size against fp64 autograd on the oracle (same
    sample positions): the numbers tests/test_gpu_split.py bounds by 1e-3 / 1e-2."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from helpers import build_model, conf_of, max_abs, rel_err, state_dict_cpu
    from oracle import volsdf_oracle as O
    import svolsdf_b200.scene as S
    res = {}
    for beta in (0.05, 0.01):
        model = build_model('dtu', perturb=True, beta=beta, device='cuda').train().set_engine(engine)
        sd = state_dict_cpu(model)
        inp, gt = S.make_input('dtu', R), S.gt_rgb(R)
        torch.manual_seed(321)
        out = model({k: v.cuda() for k, v in inp.items()}, fast=1)
        loss = TrainRig.loss_of(out, gt.cuda())
        model.zero_grad()
        loss.backward()
        ref = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
        z, z_eik = model.last_z
        torch.manual_seed(321)
        rng = O.draw_rng(R, True)
        o = O.volsdf_forward(ref, conf_of('dtu'), inp, True, fast=1, rng=rng, dtype=torch.float64,
                             z_override=(z.cpu(), z_eik.cpu(), None))
        O.volsdf_loss(o, gt).backward()
        worst = max(rel_err(p.grad.cpu(), ref[n].grad) for n, p in model.named_parameters()
                    if ref[n].grad is not None and float(ref[n].grad.norm()) > 1e-10)
        res['beta_%g' % beta] = {'rgb_max_abs': max_abs(out['rgb_values'].detach().cpu(), o['rgb_values'].detach()),
                                 'depth_max_abs': max_abs(out['depth_values'].detach().cpu(), o['depth_values'].detach()),
                                 'worst_param_grad_rel': worst}
    return {'rays': R, 'against': 'fp64 autograd on oracle/volsdf_oracle.py, same sample positions', 'errors': res,
            'bounds': {'rgb/depth max-abs': 1e-3, 'param grad rel': 1e-2}}


def gpu_eager_leg(R, steps=3):
    """the reference algorithm as plain PyTorch eager fp32 on the B200 — the implementation to beat.  The
    reference itself when /root/reference is mounted (this container has no GPU; the GPU box has no /root/reference), else
    the oracle's restatement of it, run on cuda:0 with TF32 off."""
    import svolsdf_b200.conf as C
    import svolsdf_b200.scene as S
    from oracle import volsdf_oracle as O
    from svolsdf_b200.model.network import VolSDFNetwork
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device('cuda', torch.cuda.current_device())
    torch.manual_seed(0)
    m = VolSDFNetwork(C.dtu_model_conf())
    with torch.device(dev):   # the oracle creates its constants / random draws with bare factory calls
        sd = {k: v.detach().to(dev).clone().requires_grad_(True) for k, v in m.state_dict().items()}
        opt = torch.optim.Adam(list(sd.values()), lr=5e-4)
        conf = C.dtu_model_conf()
    inp = {k: v.to(dev) for k, v in S.make_input('dtu', R).items()}
    gt = S.gt_rgb(R).to(dev)
    with torch.device(dev):

        def step():
            out = O.volsdf_forward(sd, conf, inp, True, fast=1)
            loss = O.volsdf_loss(out, gt)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(sd.values()), 1.0)
            opt.step()
            return loss
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        float(loss)
    ms = e0.elapsed_time(e1) / steps
    return {'value': R / (ms * 1e-3), 'unit': 'rays/s', 'ms_per_step': ms, 'kind': 'port',
            'what': 'oracle/volsdf_oracle.py (restatement of the reference, plain PyTorch ops) on cuda:0, fp32 eager, TF32 off, '
                    '%d rays, fwd+loss+bwd+clip+Adam, 1 warm-up + %d timed' % (R, steps)}


def frame_leg(args, engine, world, rank, dev, barrier, max_over_ranks):
    """BASELINE configs[2]: 1600x1200 DTU full-image render (rgb + depth + normal), ray-sharded over the ranks, 512-ray
    convergence groups inside every launch chunk (= the reference's chunked render loop ray for ray)."""
    import svolsdf_b200.conf as C
    import svolsdf_b200.scene as S
    from svolsdf_b200.model.network import VolSDFNetwork
    from svolsdf_b200.render import render_image
    torch.manual_seed(0)
    model = VolSDFNetwork(C.dtu_model_conf())
    beta = 0.01 if args.beta is None else args.beta
    S.perturb_(model, w_std=0.0, b_std=0.0, beta=beta)
    model = model.to(dev).eval().set_engine(engine)
    Wd, Ht = 1600, 1200
    inp = S.make_input('dtu', Wd * Ht, width=Wd, height=Ht, pixels='grid')
    K_, pose, uv = inp['intrinsics'].to(dev), inp['pose'].to(dev), inp['uv'].to(dev)
    R = uv.shape[1]
    torch.manual_seed(7)
    out = render_image(model, K_, pose, uv, chunk=args.chunk, rank=rank, world=world, group=args.group)   # warm-up
    nf = 1 if world == 1 else 2
    ms = timed(lambda i: render_image(model, K_, pose, uv, chunk=args.chunk, rank=rank, world=world, group=args.group),
               nf, barrier, max_over_ranks)
    iters = out['sampler_iters']
    mean_it = sum(iters) / max(len(iters), 1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([float(sum(iters)), float(len(iters))], device=dev)
        dist.all_reduce(t)
        mean_it = float(t[0] / t[1])
    pk = peaks()
    flop = R * (mean_it * 128 * F_SDF + 98 * (2 * F_SDF + F_RENDER))
    tf = flop / (ms * 1e-3) / 1e12
    return {'metric': 'ms/frame 1600x1200 render', 'value': ms, 'unit': 'ms/frame', 'higher_is_better': False,
            'scaling': 'strong', 'rays': R, 'rays_per_s': R / (ms * 1e-3), 'frames_timed': nf, 'beta': beta,
            'chunk_rays': args.chunk, 'convergence_group_rays': args.group, 'sampler_iterations_mean': mean_it,
            'parallelism': 'ray-sharded dp%d (512-ray groups stay on one rank) + all-gather of 28 B/ray' % world,
            'roofline': {'bound': 'tensor', 'achieved': tf / world, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                         'frac': tf / world / pk['tflops_sustained'], 'traffic': None,
                         'kernel': 'whole frame (algorithmic FLOP: I*128 F + 98 (2 F + F_r) per ray, I = mean sampler iterations)'},
            'rgb_mean': float(out['rgb_values'].mean()), 'depth_mean': float(out['depth_values'].mean())}


def mvs_leg(args, engine, dev, R, K):
    """the paper configuration's step (use_mvs: vsdf.py:205-211; loss.py:80-115 with mvs / sparsity terms):
    forward + CostMapper over 3 source views (48 x 288 x 384 probability volumes) + VolSDFLoss + backward + clip + Adam as
    one CUDA graph."""
    import svolsdf_b200.conf as C
    import svolsdf_b200.scene as S
    from svolsdf_b200.model.loss import VolSDFLoss
    from svolsdf_b200.model.network import VolSDFNetwork
    from svolsdf_b200.model.ray_sampler import RefRng, TapeRng
    from svolsdf_b200.mvs import CostMapper
    from svolsdf_b200.optim import FusedAdam
    from svolsdf_b200.train import GraphedTrainStep
    torch.manual_seed(0)
    model = VolSDFNetwork(C.dtu_model_conf()).to(dev).train().set_engine(engine)
    opt = FusedAdam(model.parameters(), lr=5e-4, max_grad_norm=1.0)
    inp = {k: v.to(dev) for k, v in S.make_input('dtu', R).items()}
    gt = {'rgb': S.gt_rgb(R).to(dev), 'rgb_smooth': S.gt_rgb(R, seed=5).to(dev)}
    views = S.mvs_views(n_views=3, dz=48, h=288, w=384, img_res=(1200, 1600), seed=9)
    ids = [25, 22, 28]
    cm = CostMapper([v['cost'][None] for v in views], [v['z_mvs'][None] for v in views], [v['K'] for v in views],
                    [v['c2w'] for v in views], ids, (1200, 1600), inverse_depth=True)
    loss_fn = VolSDFLoss(rgb_loss='torch.nn.L1Loss', eikonal_weight=0.1, mvs_weight=1.0, sparse_weight=1.0, anneal_rgb=200, gce=0.5)
    step = GraphedTrainStep(model, opt, loss_fn, inp, gt, grad_clip=0.0, cost_mapper=cm, own_view=ids[1])
    torch.manual_seed(77)
    tapes = []
    for _ in range(3 + K):
        tr = TapeRng(RefRng(dev))
        model.rng_source = tr
        with torch.no_grad():
            model(inp, fast=1)
        tapes.append(tr.tape)
    model.rng_source = step.rng
    for i in range(3):
        step(draws=tapes[i], own_view=ids[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(draws=tapes[3 + i], own_view=ids[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    return {'workload': 'MVS-supervised DTU train step (paper configuration, SURVEY.md 8f-1): forward + cost lookup in 3 source views fused with the GCE term (svs_mvs_loss) '
                        '(48x288x384 volumes) + VolSDFLoss (L1 + eikonal + GCE(0.5) on the weights + annealed sparsity) + backward '
                        '+ clip + Adam, one CUDA graph', 'rays': R, 'value': R / (ms * 1e-3), 'unit': 'rays/s', 'ms_per_step': ms,
            'steps': K, 'warmup': 3, 'loss': float(step.loss)}


MLP_KERNELS = ('mlp_',)


def run_ours(args):
    import torch.distributed as dist
    import svolsdf_b200._lib as L

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: svolsdf_b200 has no CPU path (use --impl reference)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L.load()
    engine = pick_engine(args, L)
    K, W, R = args.steps, max(args.warmup, 3), args.rays
    skip = set(filter(None, args.skip.split(',')))
    if args.workload == 'train':
        skip |= {'frame', 'large'}
    n_prof = 2
    bmvs = args.scene == 'bmvs'

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    clocks = ClockSampler(local)
    clocks.start()           # long before the timed region: nvidia-smi needs 0.1 s (1 GPU) .. 1 s (8 GPUs) to deliver its first sample
    rig = TrainRig(args, R, engine, world, rank, dev, W + max(K, n_prof), args.scene)
    Rg = rig.Rg
    step_mode = rig.step_mode
    allreduce_mode = rig.allreduce

    # ---- device-timed region: inputs resident in HBM ----
    for i in range(W):
        rig.step_device(i)
    barrier()
    clocks.mark()
    ms_step = timed(lambda i: rig.step_device(W + i), K, barrier, max_over_ranks)
    clocks.mark()
    launches = rig.launches_per_step * K
    value = Rg / (ms_step * 1e-3)

    # ---- end-to-end region: host buffers + host RNG + loss read-back every step ----
    torch.manual_seed(99 + (rank if rig.rng_mode == 'local' else 0))
    for _ in range(3):
        rig.step_e2e()
    h2d = [0]

    def e2e_step(i):
        _, nb = rig.step_e2e()
        h2d[0] = nb
    ms_e2e = timed(e2e_step, K, barrier, max_over_ranks)
    clk = clocks.stop()      # samples inside the device-timed region (or, if it was shorter than a sampling period, up to here)
    h2d_bytes = h2d[0] + sum(v.numel() * v.element_size() for v in rig.inp_pin.values()) + rig.gt_pin.numel() * 4

    # ---- per-kernel profile pass (CUDA events around every launch of the library; not part of the timing) ----
    L.prof_enable(True)
    for i in range(n_prof):
        rig.step_eager(W + i)
    torch.cuda.synchronize()
    prof = L.prof_collect()
    L.prof_enable(False)
    rig.model.rng_source = None
    pk = peaks()
    roofline, rooflines, kernels = None, [], {}
    traffic_tab = {}
    tp = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tp):
        traffic_tab = json.load(open(tp))     # per-launch dram bytes of each kernel from the committed ncu --set full capture
    if prof:
        tot_ms = sum(v['ms'] for v in prof.values())
        for name, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            kernels[name] = {'launches_per_step': v['launches'] / n_prof, 'ms_per_step': v['ms'] / n_prof,
                             'share_of_kernel_time': v['ms'] / tot_ms}

        def entry(name, v):
            """MLP kernels (dense contractions) are reported on the TENSOR roofline, sampler / compositor /
            ray kernels on the HBM roofline; the other fraction rides along.  achieved = ALGORITHMIC flops / bytes (8d
            figures, passed by the library with every launch; split-operand chains count their flops once, not 3x) over
            the live CUDA-event time."""
            sec = v['ms'] * 1e-3
            tf = v['flops'] / sec / 1e12 if v['flops'] > 0 else 0.0
            gb = v['bytes'] / sec / 1e9 if v['bytes'] > 0 else 0.0
            f_t, f_h = tf / pk['tflops_sustained'], gb / pk['hbm_gbs']
            tr = traffic_tab.get(name)
            e = {'kernel': name, 'avg_launch_ms': v['ms'] / v['launches'], 'traffic': tr,
                 'tensor_frac': f_t, 'hbm_frac': f_h if gb > 0 else None,
                 'hbm_frac_measured_traffic': (tr / (v['ms'] / v['launches'] * 1e-3) / 1e9 / pk['hbm_gbs']) if tr else None}
            if engine == L.ENGINE_TC_SPLIT and name in ('mlp_tc_sdf_fwd', 'mlp_tc_render_fwd'):
                # split-operand chains issue 3 MMAs per algorithmic one (the roofline fraction counts them once)
                e['mma_issued_frac_of_tensor_peak'] = 3.0 * f_t
            if name.startswith(MLP_KERNELS) and v['flops'] > 0:
                e.update({'bound': 'tensor', 'achieved': tf, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                          'frac': f_t, 'peak_source': pk['source'] + ' bf16 sustained'})
            else:
                e.update({'bound': 'hbm', 'achieved': gb, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': f_h,
                          'peak_source': pk['source']})
            return e
        name, v = max(prof.items(), key=lambda kv: kv[1]['ms'])
        roofline = entry(name, v)
        rooflines = [entry(n_, v_) for n_, v_ in sorted(prof.items(), key=lambda kv: -kv[1]['ms']) if v_['flops'] > 0 or v_['bytes'] > 0]

    # ---- the other engines on the same step, eager, for context (not the headline) ----
    other = {}
    if world == 1 and 'engines' not in skip:
        for ename, e in (('tc_single_fp16', L.ENGINE_TC), ('fp32_simt', L.ENGINE_FP32)):
            if e == engine:
                continue
            rig.model.set_engine(e)
            for i in range(2):
                rig.step_eager(i)
            ms_o = timed(lambda i: rig.step_eager(W + i), 3, barrier, max_over_ranks)
            other[ename] = {'ms_per_step_eager': ms_o, 'rays_per_s': Rg / (ms_o * 1e-3), 'engine': ENGINE_TEXT[e][1]}
        rig.model.set_engine(engine)
        for i in range(2):
            rig.step_eager(i)
        ms_o = timed(lambda i: rig.step_eager(W + i), 3, barrier, max_over_ranks)
        other['benchmarked_engine_eager'] = {'ms_per_step_eager': ms_o, 'rays_per_s': Rg / (ms_o * 1e-3)}
    rig.release()
    del rig
    torch.cuda.empty_cache()

    # ---- BASELINE configs[4]: large-batch data-parallel step, 8192 rays per GPU (65536 over 8 GPUs) ----
    large = None
    if 'large' not in skip and not bmvs:
        Kl = max(3, min(K, 10))
        big = TrainRig(args, args.large_rays, engine, world, rank, dev, 3 + Kl, 'dtu')
        for i in range(3):
            big.step_device(i)
        ms_big = timed(lambda i: big.step_device(3 + i), Kl, barrier, max_over_ranks)
        tfl = FLOP_PER_RAY_TRAIN * big.Rg / (ms_big * 1e-3) / 1e12
        large = {'workload': 'BASELINE configs[4]: data-parallel DTU train step, %d rays per GPU = %d rays per step over %d GPU(s), '
                             'gradient all-reduce: %s' % (args.large_rays, big.Rg, world, big.allreduce),
                 'rays_per_gpu': args.large_rays, 'global_rays': big.Rg, 'value': big.Rg / (ms_big * 1e-3), 'unit': 'rays/s',
                 'ms_per_step': ms_big, 'steps': Kl, 'warmup': 3, 'step_mode': big.step_mode, 'scaling': 'weak',
                 'algorithmic_tflops': tfl, 'algorithmic_frac_of_tensor_peak': tfl / world / pk['tflops_sustained']}
        big.release()
        del big
        torch.cuda.empty_cache()

    # ---- BASELINE configs[2]: full frame ----
    frame = None
    if 'frame' not in skip and not bmvs:
        frame = frame_leg(args, engine, world, rank, dev, barrier, max_over_ranks)

    def leave():
        """End of a multi-rank run.  Round 1 left with os._exit because destroy_process_group() blocked forever: the
        captured CUDA graphs (which hold the NCCL all-reduce nodes) were still alive when the communicator was torn down.
        Every TrainRig releases its graph before we get here (rig.release()), after which the destructor returns."""
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        dist.destroy_process_group()

    if rank != 0:
        if world > 1:
            leave()
        return
    cpu = parity = gpu_eager = mvs = None
    if world == 1 and 'mvs' not in skip and not bmvs:
        try:
            mvs = mvs_leg(args, engine, dev, R, max(3, min(K, 10)))
        except Exception as e:
            mvs = {'error': str(e).splitlines()[0][:200]}
    if world == 1:
        if not args.no_cpu_baseline:
            cpu = cpu_train_steps(args.cpu_rays, max(1, args.cpu_steps), 1)     # ~10-15 s of CPU work
            cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        if 'parity' not in skip and not bmvs:
            try:
                parity = parity_leg(engine, 1024)
            except Exception as e:
                parity = {'error': str(e).splitlines()[0][:200]}
        if 'gpu_eager' not in skip and not bmvs:
            try:
                gpu_eager = gpu_eager_leg(R)
            except Exception as e:
                gpu_eager = {'error': str(e).splitlines()[0][:200]}
    # BMVS: 128F + 97(6F+3F_r) + 32(3F_bg+3F_br) + 12F per ray
    step_tflops = (1020432384.0 if bmvs else FLOP_PER_RAY_TRAIN) * Rg / (ms_step * 1e-3) / 1e12
    line = {
        'metric': 'rays/sec (fwd+bwd train step)', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': ENGINE_TEXT[engine][0], 'data': 'synthetic',
        'config': {'workload': ('BlendedMVS VolSDFNetworkBG train step fwd+bwd+eikonal (BASELINE configs[3]): sampler 128 + main 97 + '
                                'eikonal 2 SDF evals + 32 inverted-sphere background samples/ray, L1+0.1*eik loss, clip+Adam') if bmvs else
                               ('DTU VolSDF train step fwd+bwd+eikonal (BASELINE configs[1]): sampler 128 + main 98 + '
                                'eikonal 2 SDF evals/ray, L1+0.1*eik loss, clip+Adam'), 'rays_per_gpu': R,
                   'global_rays': Rg, 'parallelism': 'ray-sharded dp%d' % world,
                   'l2': 'per-step working set (~2 GB of saved activation tiles) exceeds the 126 MB L2; no explicit flush',
                   'engine': ENGINE_TEXT[engine][1],
                   'step_mode': step_mode, 'gradient_allreduce': allreduce_mode,
                   'host_rng': {'single': 'reference order, CPU default generator', 'local': 'per-rank CPU streams (seed + rank)',
                                'global': 'global-batch draws on every rank, rows sliced (sharded == unsharded per ray)'}[
                                    'single' if world == 1 else ('global' if args.rng == 'global' else 'local')]},
        'e2e': {'value': Rg / (ms_e2e * 1e-3), 'unit': 'rays/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches), 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
        'algorithmic_tflops': step_tflops, 'algorithmic_frac_of_tensor_peak': step_tflops * 1.0 / world / pk['tflops_sustained'],
        'parity': parity, 'gpu_eager_baseline': gpu_eager, 'frame': frame, 'large_batch': large, 'mvs_step': mvs,
        'kernels': kernels, 'rooflines': rooflines, 'other_engines': other,
    }
    print(json.dumps(line))
    if world > 1:
        leave()


F_SDF, F_RENDER = 1049088.0, 533504.0   # FLOP per point-forward, SURVEY.md Appendix B


def run_frame(args):
    """`--workload frame`: only BASELINE configs[2] (the default run carries the same numbers under `frame`)."""
    import torch.distributed as dist
    import svolsdf_b200._lib as L
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    L.load()
    engine = pick_engine(args, L)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms
    clocks = ClockSampler(local)
    clocks.start()
    fr = frame_leg(args, engine, world, rank, dev, barrier, max_over_ranks)
    clk = clocks.stop()
    if rank == 0:
        fr.update({'n_gpus': world, 'steps': fr['frames_timed'], 'warmup': 1, 'ms_per_step': fr['value'], 'vs_baseline': None,
                   'dtype': ENGINE_TEXT[engine][0], 'data': 'synthetic', 'clocks': clk,
                   'config': {'workload': 'DTU 1600x1200 full-image render (BASELINE configs[2]): rgb + depth + normal',
                              'engine': ENGINE_TEXT[engine][1]}})
        print(json.dumps(fr))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    elif a.workload == 'frame':
        run_frame(a)
    else:
        run_ours(a)
