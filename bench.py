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
    ap.add_argument('--engine', default='auto', choices=['auto', 'fp32', 'tc'])
    ap.add_argument('--cpu-rays', type=int, default=128, help='rays per step of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='train', choices=['train', 'frame'],
                    help="train: BASELINE configs[1] (default, the headline); frame: configs[2], 1600x1200 full-image render")
    ap.add_argument('--chunk', type=int, default=16384, help='frame workload: rays per model call (= sampler convergence group)')
    ap.add_argument('--beta', type=float, default=None, help='frame workload: density.beta override (0.01 = trained-like, 5 sampler iterations)')
    ap.add_argument('--no-parity-leg', action='store_true', help='skip the fp32 parity-engine comparison leg (profiling runs)')
    ap.add_argument('--scene', default='dtu', choices=['dtu', 'bmvs'],
                    help='train workload: dtu = BASELINE configs[1] (the headline); bmvs = configs[3], VolSDFNetworkBG with the '
                         'inverted-sphere background networks at 768x576')
    ap.add_argument('--rng', default='auto', choices=['auto', 'global', 'local'],
                    help='multi-GPU host random draws: global = every rank draws the whole batch and keeps its rows (sharded run '
                         'bit-identical per ray to the unsharded one, host work grows with N); local = per-rank streams (data-parallel '
                         'semantics, host work constant); auto = local when N > 1')
    ap.add_argument('--eager', action='store_true', help='issue every step from Python instead of replaying a CUDA graph')
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

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
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
                'samples': len(sm)}


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
    n = max(32, min(args.cpu_rays * 2, 256))
    steps = max(1, min(args.steps, 8))
    r = cpu_train_steps(n, steps, min(args.warmup, 1))
    line = {
        'metric': 'rays/sec (fwd+bwd train step)', 'value': r['value'], 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': 'DTU VolSDF train step (BASELINE configs[1]) on host CPU cores, bounded sample of %d rays/step' % n,
                   'rays_per_step': n},
        'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': r['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch.distributed as dist
    import svolsdf_b200._lib as L
    import svolsdf_b200.conf as C
    import svolsdf_b200.scene as S
    from svolsdf_b200 import dist as sdist
    from svolsdf_b200.model.network import VolSDFNetwork
    from svolsdf_b200.model.ray_sampler import RecordedRng, RefRng, TapeRng

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
    engine = L.ENGINE_FP32
    if args.engine == 'tc' or (args.engine == 'auto' and L.load().svs_has_engine(L.ENGINE_TC)):
        engine = L.ENGINE_TC
    K, W, R = args.steps, max(args.warmup, 3), args.rays
    Rg = R * world

    torch.manual_seed(0)
    bmvs = args.scene == 'bmvs'
    if bmvs:
        from svolsdf_b200.model.network_bg import VolSDFNetworkBG
        model = VolSDFNetworkBG(C.bmvs_model_conf()).to(dev).train().set_engine(engine)
    else:
        model = VolSDFNetwork(C.dtu_model_conf()).to(dev).train().set_engine(engine)
    # clip_grad_norm_(1.0) + NaN guard + Adam of vsdf.py:214-219 as one fused step (svolsdf_b200.optim.FusedAdam)
    from svolsdf_b200.optim import FusedAdam
    opt = FusedAdam(model.parameters(), lr=5e-4, max_grad_norm=1.0)
    reducer = sdist.GradAllReducer(model.parameters()) if world > 1 else None
    inp_host = S.make_input(args.scene, Rg, pixels='perm' if Rg > 4096 else 'random')
    gt_host = S.gt_rgb(Rg)
    lo, hi = sdist.shard_range(Rg, rank, world)
    inp_host = sdist.shard_input(inp_host, rank, world)
    gt_host = gt_host[:, lo:hi].contiguous()
    inp_pin = {k: v.pin_memory() for k, v in inp_host.items()}
    gt_pin = gt_host.pin_memory()
    inp_dev = {k: v.to(dev) for k, v in inp_host.items()}
    gt_dev = gt_host.to(dev)

    def loss_of(out, gt):
        return (out['rgb_values'] - gt.reshape(-1, 3)).abs().mean() + \
            0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()

    def finish(loss):
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if reducer is not None:
            reducer.allreduce_(world)
        opt.step()      # gradient clipping to norm 1.0 happens inside the fused step

    rng_mode = 'single' if world == 1 else ('global' if args.rng == 'global' else 'local')

    def make_rng():
        return sdist.ShardedRng(dev, Rg, lo, hi) if rng_mode == 'global' else RefRng(dev)

    # random draws of every step, made in the reference's order and uploaded BEFORE the timed region
    torch.manual_seed(1234 + (rank if rng_mode == 'local' else 0))
    tapes = []
    n_prof = 2
    for _ in range(W + max(K, n_prof)):
        tr = TapeRng(make_rng())
        model.rng_source = tr
        with torch.no_grad():   # a dry forward only to make the draws in the reference's order (not timed)
            model(inp_dev, fast=1)
        tapes.append(tr.tape)
    model.rng_source = None
    torch.cuda.synchronize()

    def step_eager(i):
        model.rng_source = RecordedRng(dev, tapes[i])
        out = model(inp_dev, fast=1)
        finish(loss_of(out, gt_dev))

    # launches of this library per step (counted on one eager step; a graph replay re-issues exactly these)
    l0 = L.launch_count()
    step_eager(0)
    torch.cuda.synchronize()
    launches_per_step = L.launch_count() - l0

    graphed, step_mode = None, 'eager'
    if not args.eager:
        try:
            from svolsdf_b200.train import GraphedTrainStep
            graphed = GraphedTrainStep(model, opt, loss_of, inp_dev, gt_dev, grad_clip=0.0, reducer=reducer, world=world,
                                       make_rng=make_rng)
            step_mode = 'cuda_graph'
        except Exception as e:   # keep the run alive, say so in the JSON line
            graphed, step_mode = None, 'eager (graph capture failed: %s)' % (str(e).splitlines()[0][:120],)
            model.rng_source = None

    def step_device(i):
        if graphed is not None:
            graphed(draws=tapes[i])     # inputs + this step's random draws already resident in HBM
        else:
            step_eager(i)

    pending = {'draws': None}

    def step_e2e():
        if graphed is not None:
            # host RNG in the reference's order -> pinned -> device.  The draws of step i+1 are made on the CPU while
            # the GPU replays step i (they only depend on the CPU generator), then the loss of step i is read back.
            draws = pending['draws'] if pending['draws'] is not None else graphed.draw()
            loss = graphed(inp_pin, gt_pin, draws)
            pending['draws'] = graphed.draw()
            return float(loss.item()), graphed.h2d_bytes_rng
        model.rng_source = make_rng()
        inp = {k: v.to(dev, non_blocking=True) for k, v in inp_pin.items()}
        gt = gt_pin.to(dev, non_blocking=True)
        out = model(inp, fast=1)
        loss = loss_of(out, gt)
        finish(loss)
        return float(loss.item()), model.rng_source.h2d_bytes   # device->host read of the step's result

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

    # ---- device-timed region: inputs resident in HBM ----
    for i in range(W):
        step_device(i)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step_device(W + i)
    e1.record()
    barrier()
    launches = launches_per_step * K
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop()
    ms_step = ms_total / K
    value = Rg / (ms_step * 1e-3)

    # ---- end-to-end region: host buffers + host RNG + loss read-back every step ----
    torch.manual_seed(99 + (rank if rng_mode == 'local' else 0))
    for _ in range(3):
        step_e2e()
    barrier()
    h2d = 0
    e0.record()
    for _ in range(K):
        _, nb = step_e2e()
        h2d = nb
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / K
    h2d_bytes = h2d + sum(v.numel() * v.element_size() for v in inp_pin.values()) + gt_pin.numel() * 4

    # ---- per-kernel profile pass (CUDA events around every launch of the library; not part of the timing) ----
    L.prof_enable(True)
    for i in range(n_prof):
        step_eager(W + i)
    torch.cuda.synchronize()
    prof = L.prof_collect()
    L.prof_enable(False)
    model.rng_source = None
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
            """both rooflines of one kernel; `bound` = the one it sits closer to.  achieved = ALGORITHMIC flops / bytes
            (SURVEY.md 8d figures, passed by the library with every launch) over the CUDA-event time."""
            sec = v['ms'] * 1e-3
            tf = v['flops'] / sec / 1e12 if v['flops'] > 0 else 0.0
            gb = v['bytes'] / sec / 1e9 if v['bytes'] > 0 else 0.0
            f_t, f_h = tf / pk['tflops_sustained'], gb / pk['hbm_gbs']
            tr = traffic_tab.get(name)
            e = {'kernel': name, 'avg_launch_ms': v['ms'] / v['launches'], 'traffic': tr,
                 'tensor_frac': f_t, 'hbm_frac': f_h if gb > 0 else None,
                 # measured DRAM bytes (ncu, per launch) over the live launch time: how close the kernel is to the HBM
                 # roofline counting everything it actually moves (saved activation tiles included)
                 'hbm_frac_measured_traffic': (tr / (v['ms'] / v['launches'] * 1e-3) / 1e9 / pk['hbm_gbs']) if tr else None}
            if f_h > f_t:
                e.update({'bound': 'hbm', 'achieved': gb, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': f_h,
                          'peak_source': pk['source']})
            else:
                e.update({'bound': 'tensor', 'achieved': tf, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                          'frac': f_t, 'peak_source': pk['source'] + ' bf16 sustained'})
            return e
        name, v = max(prof.items(), key=lambda kv: kv[1]['ms'])
        roofline = entry(name, v)
        rooflines = [entry(n_, v_) for n_, v_ in sorted(prof.items(), key=lambda kv: -kv[1]['ms']) if v_['flops'] > 0 or v_['bytes'] > 0]

    # ---- the fp32 parity engine on the same step, for reference (not the headline) ----
    parity = None
    if engine != L.ENGINE_FP32 and world == 1 and not args.no_parity_leg:
        model.set_engine(L.ENGINE_FP32)
        for i in range(2):
            step_eager(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(3):
            step_eager(W + i % max(K, n_prof))
        e1.record()
        torch.cuda.synchronize()
        parity = {'engine': 'fp32 SIMT (parity mode)', 'ms_per_step': e0.elapsed_time(e1) / 3,
                  'rays_per_s': Rg / (e0.elapsed_time(e1) / 3 * 1e-3)}
        model.set_engine(engine)
    model.rng_source = None

    def leave():
        """End of a multi-rank run.  destroy_process_group() was seen to block forever once the communicator has been
        used inside a captured CUDA graph (2-GPU run: the JSON line was out, both ranks sat in the destructor until the
        launcher's timeout): drop the graph, drain the device, meet once more, then leave without the destructor."""
        nonlocal graphed
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    if rank != 0:
        if world > 1:
            leave()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_train_steps(args.cpu_rays, 2, 1)
        cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    # BMVS: 128F + 97(6F+3F_r) + 32(3F_bg+3F_br) + 12F per ray
    step_tflops = (1020432384.0 if bmvs else FLOP_PER_RAY_TRAIN) * Rg / (ms_step * 1e-3) / 1e12
    line = {
        'metric': 'rays/sec (fwd+bwd train step)', 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f16' if engine == L.ENGINE_TC else 'f32', 'data': 'synthetic',
        'config': {'workload': ('BlendedMVS VolSDFNetworkBG train step fwd+bwd+eikonal (BASELINE configs[3]): sampler 128 + main 97 + '
                                'eikonal 2 SDF evals + 32 inverted-sphere background samples/ray, L1+0.1*eik loss, clip+Adam') if bmvs else
                               ('DTU VolSDF train step fwd+bwd+eikonal (BASELINE configs[1]): sampler 128 + main 98 + '
                                'eikonal 2 SDF evals/ray, L1+0.1*eik loss, clip+Adam'), 'rays_per_gpu': R,
                   'global_rays': Rg, 'parallelism': 'ray-sharded dp%d' % world,
                   'l2': 'per-step working set (~2 GB of saved activation tiles) exceeds the 126 MB L2; no explicit flush',
                   'engine': 'tcgen05 kind::f16 (fp16 operands, fp32 TMEM accumulate)' if engine == L.ENGINE_TC else 'fp32 SIMT (parity mode)',
                   'step_mode': step_mode,
                   'host_rng': {'single': 'reference order, CPU default generator', 'local': 'per-rank CPU streams (seed + rank)',
                                'global': 'global-batch draws on every rank, rows sliced (sharded == unsharded per ray)'}[rng_mode]},
        'e2e': {'value': Rg / (ms_e2e * 1e-3), 'unit': 'rays/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches), 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
        'algorithmic_tflops': step_tflops, 'algorithmic_frac_of_tensor_peak': step_tflops * 1.0 / world / pk['tflops_sustained'],
        'kernels': kernels, 'rooflines': rooflines, 'parity_engine': parity,
    }
    print(json.dumps(line))
    if world > 1:
        leave()


F_SDF, F_RENDER = 1049088.0, 533504.0   # FLOP per point-forward, SURVEY.md Appendix B


def run_frame(args):
    """BASELINE configs[2]: 1600x1200 DTU full-image render (rgb + depth + normal), ray-sharded over the ranks."""
    import torch.distributed as dist
    import svolsdf_b200._lib as L
    import svolsdf_b200.conf as C
    import svolsdf_b200.scene as S
    from svolsdf_b200.model.network import VolSDFNetwork
    from svolsdf_b200.render import render_image

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    engine = L.ENGINE_FP32 if args.engine == 'fp32' else L.ENGINE_TC
    torch.manual_seed(0)
    model = VolSDFNetwork(C.dtu_model_conf())
    if args.beta is not None:
        S.perturb_(model, w_std=0.0, b_std=0.0, beta=args.beta)
    model = model.to(dev).eval().set_engine(engine)
    Wd, Ht = 1600, 1200
    inp = S.make_input('dtu', Wd * Ht, width=Wd, height=Ht, pixels='grid')
    K_, pose, uv = inp['intrinsics'].to(dev), inp['pose'].to(dev), inp['uv'].to(dev)
    R = uv.shape[1]
    K, W = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(7)
    for _ in range(W):
        out = render_image(model, K_, pose, uv, chunk=args.chunk, rank=rank, world=world)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        out = render_image(model, K_, pose, uv, chunk=args.chunk, rank=rank, world=world)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / K
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop()
    iters = out['sampler_iters']
    if rank == 0:
        pk = peaks()
        mean_it = sum(iters) / max(len(iters), 1)
        flop = R * (mean_it * 128 * F_SDF + 98 * (2 * F_SDF + F_RENDER))
        tf = flop / (ms * 1e-3) / 1e12
        print(json.dumps({
            'metric': 'ms/frame 1600x1200 render', 'value': ms, 'unit': 'ms/frame', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms, 'higher_is_better': False, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f16' if engine == L.ENGINE_TC else 'f32', 'data': 'synthetic',
            'config': {'workload': 'DTU 1600x1200 full-image render (BASELINE configs[2]): rgb + depth + normal, %d rays, '
                                   'chunks of %d rays per model call' % (R, args.chunk), 'beta': args.beta,
                       'sampler_iterations_mean': mean_it, 'parallelism': 'ray-sharded dp%d + all-gather of 28 B/ray' % world},
            'rays_per_s': R / (ms * 1e-3), 'clocks': clk,
            'roofline': {'bound': 'tensor', 'achieved': tf / world, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                         'frac': tf / world / pk['tflops_sustained'], 'traffic': None, 'kernel': 'whole frame (algorithmic FLOP)'},
            'rgb_mean': float(out['rgb_values'].mean()), 'depth_mean': float(out['depth_values'].mean())}))
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
