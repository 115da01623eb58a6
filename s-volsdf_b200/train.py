"""One VolSDF optimisation step (`VolOpt.train_step`, volsdf/vsdf.py:196-219) as a replayable CUDA graph.

At 1024 rays a step is ~40 launches of this library plus ~150 small PyTorch launches (loss, clip, Adam); issuing
them from Python takes about as long as the GPU needs to run them.  Training has no data-dependent control flow
(`fast=1`: exactly one sampler iteration, ray_sampler.py:68,83), so forward + loss + backward (+ gradient all-reduce) +
clip + Adam are captured ONCE into a `torch.cuda.CUDAGraph`; every later step is: copy the step's inputs and the
reference's CPU random draws into static device buffers, replay.

    step = GraphedTrainStep(model, optimizer, loss_fn, example_input, example_gt)
    loss = step(model_input, gt_rgb)          # device scalar; .item() it when you need the number
"""
import torch

from .model.ray_sampler import RecordedRng, RefRng, TapeRng


class _StaticRng(RecordedRng):
    """Hands out the SAME device tensors every step (graph replays read them); `refill` copies fresh draws in."""

    def __init__(self, device, tape):
        super().__init__(device, tape)

    def rewind(self):
        self.pos = 0


def default_loss(out, gt):
    """L1 rgb + 0.1 * eikonal (config/vol/dtu.yaml:16-20 -> volsdf/model/loss.py:38-51)"""
    return (out['rgb_values'] - gt.reshape(-1, 3)).abs().mean() + \
        0.1 * ((out['grad_theta'].norm(2, dim=1) - 1) ** 2).mean()


class GraphedTrainStep(object):
    """loss_fn: either `f(model_outputs, gt) -> scalar` (rgb + eikonal, the default), or a `VolSDFLoss` module — then the
    step is the MVS-supervised one of the paper configuration (vsdf.py:205-211 with `use_mvs`): `cost_mapper` (a
    svolsdf_b200.mvs.CostMapper) looks up p_j / p_i for the step's samples inside the graph, `example_gt` is the
    reference's ground-truth dict ({'rgb', 'rgb_smooth'}), the batch's image index lives in `self.own_view` (device int32)
    and the iteration counter that drives the annealing in `self.iter_step` (device float, incremented by the graph).
    Optimizers other than FusedAdam get no NaN guard inside the graph (on_after_backward needs a host sync)."""

    def __init__(self, model, optimizer, loss_fn=default_loss, example_input=None, example_gt=None, grad_clip=1.0,
                 reducer=None, world=1, make_rng=None, warmup=3, cost_mapper=None, own_view=0, iter_step=0, fused_mvs=True):
        assert example_input is not None and example_gt is not None
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.grad_clip, self.reducer, self.world = grad_clip, reducer, world
        self.cost_mapper = cost_mapper
        self.fused_mvs = bool(fused_mvs)
        self.module_loss = hasattr(loss_fn, 'forward_device')
        if getattr(loss_fn, 'iter_step', None) is not None and not self.module_loss:
            raise ValueError('loss modules with a host-side step counter cannot be captured (their annealing would freeze); '
                             'pass svolsdf_b200.model.loss.VolSDFLoss or a stateless function')
        if isinstance(example_gt, dict):
            dev = next(iter(example_gt.values())).device
        else:
            dev = example_gt.device
        self.device = dev
        self.make_rng = make_rng or (lambda: RefRng(dev))
        # static inputs
        self.inp = {k: v.detach().clone() for k, v in example_input.items() if torch.is_tensor(v)}
        self.gt = ({k: v.detach().clone() for k, v in example_gt.items()} if isinstance(example_gt, dict)
                   else example_gt.detach().clone())
        self.own_view = torch.tensor([int(own_view)], dtype=torch.int32, device=dev)
        self.iter_step = torch.tensor(float(iter_step), dtype=torch.float32, device=dev)
        # one eager dry run records which random tensors a step draws (shapes / order of the reference)
        tape = TapeRng(self.make_rng())
        model.rng_source = tape
        with torch.no_grad():
            model(self.inp, fast=1)
        self.rng = _StaticRng(dev, [t.clone() for t in tape.tape])
        self.calls = list(tape.calls)
        self.h2d_bytes_rng = sum(t.numel() * t.element_size() for t in self.rng.tape)
        model.rng_source = self.rng

        # warm-up + capture run real optimisation steps: remember parameters / optimizer state and put them back
        # afterwards (in place, the graph holds their addresses)
        saved_iter = self.iter_step.clone()
        saved_params = [p.detach().clone() for p in model.parameters()]
        saved_state = [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in self.opt.state.get(p, {}).items()}
                       for p in model.parameters()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        torch.cuda.synchronize()
        with torch.no_grad():
            for p, sp, st in zip(model.parameters(), saved_params, saved_state):
                p.copy_(sp)
                for k, v in self.opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        v.copy_(st[k]) if k in st else v.zero_()
            self.iter_step.copy_(saved_iter)
        torch.cuda.synchronize()

    def _body(self):
        self.rng.rewind()
        out = self.model(self.inp, fast=1)
        if self.cost_mapper is not None:   # vsdf.py:207-209
            if self.module_loss and getattr(self.loss_fn, 'mvs_weight', 0) > 0 and self.fused_mvs:
                # lookup + GCE term in one kernel: p_i p_j stay in registers
                out['mvs_loss_fused'], out['conf_ray'] = self.cost_mapper.mvs_loss(
                    out['weights'], self.own_view, out['xyz'], gce=self.loss_fn.gce, confi=self.loss_fn.confi)
            else:
                out['pj'], out['pi'], _ = self.cost_mapper(out['depth_vals'], self.own_view, out['xyz'])
        if self.module_loss:
            loss = self.loss_fn.forward_device(out, self.gt, self.iter_step)['loss']
            self.iter_step.add_(1.0)
        else:
            loss = self.loss_fn(out, self.gt)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.allreduce_(self.world)
        if self.grad_clip:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.grad_clip)
        self.opt.step()
        return loss.detach()

    def draw(self, source=None):
        """Makes this step's random draws in the reference's order (same calls the model made in the dry run) and
        returns them as device tensors (uploaded through pinned memory by RefRng)."""
        src = source or self.make_rng()
        return [getattr(src, name)(*args) for name, args in self.calls]

    def load_draws(self, tape):
        """tape: list of device (or pinned host) tensors in draw order"""
        for dst, src in zip(self.rng.tape, tape):
            dst.copy_(src, non_blocking=True)

    def __call__(self, model_input=None, gt=None, draws=None, own_view=None):
        if model_input is not None:
            for k, v in self.inp.items():
                v.copy_(model_input[k], non_blocking=True)
        if gt is not None:
            if isinstance(self.gt, dict):
                for k, v in self.gt.items():
                    v.copy_(gt[k], non_blocking=True)
            else:
                self.gt.copy_(gt, non_blocking=True)
        if own_view is not None:
            self.own_view.fill_(int(own_view))
        if draws is not None:
            self.load_draws(draws)
        self.graph.replay()
        return self.loss
