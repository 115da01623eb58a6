"""FusedAdam (clip + NaN guard + Adam in two launches) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam."""
import pytest
import torch

from svolsdf_b200.optim import FusedAdam

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(256, 39), (256,), (256, 1), (217, 256), (257, 256), (3, 256), (), (5000,), (1,)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]


@pytest.mark.parametrize('max_norm', [0.0, 1.0])
def test_matches_torch_adam_and_clip(max_norm):
    a, b = _params(0), _params(0)
    ref = torch.optim.Adam(a, lr=5e-4)
    fus = FusedAdam(b, lr=5e-4, max_grad_norm=max_norm)
    g = torch.Generator().manual_seed(1)
    for it in range(4):
        scale = 10.0 if it % 2 == 0 else 1e-3     # above and below the clipping threshold
        for pa, pb in zip(a, b):
            gr = (torch.randn(pa.shape, generator=g) * scale).to(DEV)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        if max_norm > 0:
            n_ref = torch.nn.utils.clip_grad_norm_(a, max_norm)
        ref.step()
        fus.step()
        if max_norm > 0:
            assert abs(float(fus.last_grad_norm_sq.sqrt()) - float(n_ref)) < 1e-4 * float(n_ref)
        for pa, pb in zip(a, b):
            assert torch.allclose(pa, pb, rtol=1e-5, atol=2e-7), it
    sa, sb = ref.state_dict(), fus.state_dict()
    assert set(sa['state'][0].keys()) == set(sb['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
    assert float(sb['state'][0]['step']) == 4.0


def test_nonfinite_gradients_are_zeroed_but_adam_still_steps():
    """volsdf/vsdf.py:454-464: a NaN/Inf gradient zeroes all gradients of the step; Adam.step() still runs"""
    a, b = _params(2), _params(2)
    ref = torch.optim.Adam(a, lr=5e-4)
    fus = FusedAdam(b, lr=5e-4, max_grad_norm=1.0)
    for it in range(2):
        for i, (pa, pb) in enumerate(zip(a, b)):
            gr = torch.ones_like(pa) * 0.01
            pa.grad, pb.grad = gr.clone(), gr.clone()
        if it == 1:
            b[3].grad[5, 7] = float('nan')
            for pa in a:
                pa.grad.zero_()
        else:
            torch.nn.utils.clip_grad_norm_(a, 1.0)
        ref.step()
        fus.step()
        for pa, pb in zip(a, b):
            assert torch.isfinite(pb).all()
            assert torch.allclose(pa, pb, rtol=1e-5, atol=2e-7)


def test_resume_from_a_plain_torch_adam_checkpoint():
    """The reference saves `torch.optim.Adam(...).state_dict()` (vsdf.py checkpoints): no `capturable`, `step` an int
    (torch 1.9) or a CPU tensor.  FusedAdam.load_state_dict must turn that into device-resident float32 steps before the
    kernel sees a pointer (a host pointer used to reach svs_adam_step)."""
    a, b = _params(5), _params(5)
    ref = torch.optim.Adam(a, lr=5e-4)
    g = torch.Generator().manual_seed(6)
    grads = [[(torch.randn(p.shape, generator=g) * 0.1).to(DEV) for p in a] for _ in range(4)]
    for it in range(2):
        for pa, gr in zip(a, grads[it]):
            pa.grad = gr.clone()
        ref.step()
    import copy
    sd = copy.deepcopy(ref.state_dict())      # (state_dict() hands out the live per-parameter dicts)
    for st in sd['state'].values():            # what torch 1.9 wrote: a Python int
        st['step'] = int(st['step'])
    sd['param_groups'][0].pop('capturable', None)
    with torch.no_grad():
        for pa, pb in zip(a, b):
            pb.copy_(pa)
    fus = FusedAdam(b, lr=5e-4, max_grad_norm=0.0)
    fus.load_state_dict(sd)
    assert fus.param_groups[0]['capturable'] is True
    assert all(s['step'].is_cuda and s['step'].dtype == torch.float32 and float(s['step']) == 2.0 for s in fus.state.values())
    for it in range(2, 4):
        for pa, pb, gr in zip(a, b, grads[it]):
            pa.grad, pb.grad = gr.clone(), gr.clone()
        ref.step()
        fus.step()
    for pa, pb in zip(a, b):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=2e-7)


def test_peer_allreduce_step_equals_the_two_launch_step_on_one_gpu():
    """svs_adam_step_allreduce with world = 1 (the multi-GPU kernel reading its own buffer) == svs_adam_step; the N > 1
    path is exercised by tools/peer_adam_check.py under torchrun (replicas bit-identical, == NCCL all-reduce + step)."""
    from svolsdf_b200.dist import PeerGradBuffer
    a, b = _params(8), _params(8)
    ref = FusedAdam(a, lr=5e-4, max_grad_norm=1.0)
    peer = PeerGradBuffer(b)
    assert peer.world == 1 and peer.n_flat % 4 == 0 and all(o % 4 == 0 for o in peer.offsets)
    fus = FusedAdam(b, lr=5e-4, max_grad_norm=1.0, peer=peer)
    g = torch.Generator().manual_seed(9)
    for it in range(5):
        scale = 10.0 if it % 2 == 0 else 1e-3
        for pa, pb in zip(a, b):
            gr = (torch.randn(pa.shape, generator=g) * scale).to(DEV)
            if it == 3 and pa.dim() == 2 and pa.shape[0] == 217:
                gr[5, 7] = float('inf')
            pa.grad, pb.grad = gr.clone(), gr.clone()
        ref.step()
        fus.step()
        for pa, pb in zip(a, b):
            assert torch.isfinite(pb).all()
            assert torch.allclose(pa, pb, rtol=1e-5, atol=3e-7), it
