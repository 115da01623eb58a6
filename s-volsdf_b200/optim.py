"""`FusedAdam`: torch.optim.Adam whose step() is ONE call into the library — gradient clipping by global norm, the
reference's NaN/Inf guard and the Adam update for all parameter tensors in two launches
(volsdf/vsdf.py:214-219,454-464 issue clip_grad_norm_, one isnan/isinf host sync per parameter, and Adam.step).

State layout (`exp_avg`, `exp_avg_sq`, `step`) and `state_dict()` are those of `torch.optim.Adam`; `load_state_dict`
accepts the reference's checkpoints (plain `torch.optim.Adam` state: `step` an int (torch 1.9) or a CPU tensor,
`capturable` absent or False) and normalises them to what the kernel needs — `step` as a float32 scalar on the
parameter's device, `capturable=True`.  CUDA-graph safe.

Guard semantics (volsdf/vsdf.py:454-464 under the reference's pinned torch 1.9): non-finite gradients are zeroed and Adam
still steps (moments decay, the step count advances).  All tensors of a group must receive a gradient in the same steps
(the kernel takes one step count per group; the VolSDF models always do).
"""
import ctypes as C

import torch

from . import _lib as L


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=0.0, skip_nonfinite=True, peer=None):
        super().__init__(params, lr=lr, betas=betas, eps=eps, capturable=True, foreach=False)
        self.max_grad_norm = float(max_grad_norm)
        self.skip_nonfinite = bool(skip_nonfinite)
        # svolsdf_b200.dist.PeerGradBuffer: data-parallel step = all-reduce(mean) over NVLink peer memory + clip + guard +
        # Adam in ONE kernel (svs_adam_step_allreduce); a single parameter group holding exactly peer.params
        self.peer = peer
        self._scratch = None
        self.last_grad_norm_sq = None   # device tensor: sum of squared gradients of the last step (before clipping)

    def _step_peer(self):
        peer = self.peer
        if len(self.param_groups) != 1 or [id(p) for p in self.param_groups[0]['params'] if p.requires_grad] != [id(p) for p in peer.params]:
            raise L.SvsError('FusedAdam(peer=...): one parameter group holding exactly the PeerGradBuffer parameters')
        group = self.param_groups[0]
        ps = peer.params
        dev = ps[0].device
        for p in ps:
            st = self.state[p]
            if len(st) == 0:
                st['step'] = torch.zeros((), dtype=torch.float32, device=dev)
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if not p.is_contiguous():
                raise L.SvsError('FusedAdam needs contiguous parameters')
        if self._scratch is None or self._scratch.device != dev:
            self._scratch = torch.zeros(2, dtype=torch.float32, device=dev)
        steps = [self.state[p]['step'] for p in ps]
        if not all(t.is_cuda and t.dtype == torch.float32 for t in steps):
            self._normalise_state()
            steps = [self.state[p]['step'] for p in ps]
        peer.load_grads()
        torch._foreach_add_(steps, 1.0)
        n = len(ps)
        arr, i64 = C.c_void_p * n, C.c_int64 * n
        parr = C.c_void_p * peer.world
        L.call('svs_adam_step_allreduce', n, arr(*[p.data_ptr() for p in ps]),
               arr(*[self.state[p]['exp_avg'].data_ptr() for p in ps]), arr(*[self.state[p]['exp_avg_sq'].data_ptr() for p in ps]),
               i64(*[p.numel() for p in ps]), peer.world, peer.rank, parr(*peer.peer_grads),
               parr(*peer.peer_flags) if peer.peer_flags else None, peer.n_flat, peer.gmean.data_ptr(),
               float(group['lr']), float(group['betas'][0]), float(group['betas'][1]), float(group['eps']),
               self.max_grad_norm, 1 if self.skip_nonfinite else 0, steps[0].data_ptr(), self._scratch.data_ptr(),
               peer.ctrl.data_ptr(), L.stream())
        self.last_grad_norm_sq = self._scratch[0]
        from . import functional as F_
        F_.weights_changed()

    def _normalise_state(self):
        for group in self.param_groups:
            group['capturable'] = True
            group['foreach'] = False
            group['fused'] = None
            for p in group['params']:
                st = self.state.get(p)
                if not st:
                    continue
                step = st.get('step', 0)
                step = float(step.item()) if torch.is_tensor(step) else float(step)
                st['step'] = torch.full((), step, dtype=torch.float32, device=p.device)
                for k in ('exp_avg', 'exp_avg_sq'):
                    if k in st:
                        st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._normalise_state()

    def __setstate__(self, state):
        super().__setstate__(state)
        self._normalise_state()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.peer is not None:
            self._step_peer()
            return loss
        for group in self.param_groups:
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            if len(ps) > 96:
                raise L.SvsError('FusedAdam: more than 96 tensors in a parameter group')
            dev = ps[0].device
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = torch.zeros((), dtype=torch.float32, device=dev)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise L.SvsError('FusedAdam needs contiguous parameters and gradients')
            if self._scratch is None or self._scratch.device != dev:
                self._scratch = torch.zeros(2, dtype=torch.float32, device=dev)
            steps = [self.state[p]['step'] for p in ps]
            if not all(t.is_cuda and t.dtype == torch.float32 for t in steps):
                self._normalise_state()          # state injected behind load_state_dict's back
                steps = [self.state[p]['step'] for p in ps]
            torch._foreach_add_(steps, 1.0)      # every tensor of a group carries the same count, as in torch
            n = len(ps)
            arr = C.c_void_p * n
            i64 = C.c_int64 * n
            L.call('svs_adam_step', n, arr(*[p.data_ptr() for p in ps]), arr(*[p.grad.data_ptr() for p in ps]),
                   arr(*[self.state[p]['exp_avg'].data_ptr() for p in ps]),
                   arr(*[self.state[p]['exp_avg_sq'].data_ptr() for p in ps]), i64(*[p.numel() for p in ps]),
                   float(group['lr']), float(group['betas'][0]), float(group['betas'][1]), float(group['eps']),
                   self.max_grad_norm, 1 if self.skip_nonfinite else 0, steps[0].data_ptr(), self._scratch.data_ptr(),
                   L.stream())
            self.last_grad_norm_sq = self._scratch[0]
        from . import functional as F_
        F_.weights_changed()      # parameters were written through raw pointers: invalidate the packed-weight caches
        return loss
