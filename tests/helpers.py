"""Shared test helpers: model construction (seeded like the goldens) and golden loading."""
import os
import warnings

import numpy as np
import torch

import svolsdf_b200.conf as C
import svolsdf_b200.scene as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
warnings.filterwarnings('ignore', message='.*weight_norm.*')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def build_model(kind='dtu', perturb=False, beta=None, device='cpu'):
    """Same recipe as oracle/make_golden.py::build, with OUR model classes."""
    from svolsdf_b200.model.network import VolSDFNetwork
    from svolsdf_b200.model.network_bg import VolSDFNetworkBG
    torch.manual_seed(0)
    model = VolSDFNetwork(C.dtu_model_conf()) if kind == 'dtu' else VolSDFNetworkBG(C.bmvs_model_conf())
    if perturb or beta is not None:
        S.perturb_(model, seed=7, w_std=S.PERTURB_W if perturb else 0.0, b_std=S.PERTURB_B if perturb else 0.0, beta=beta)
    return model.to(device)


def model_from_golden(g, device='cpu'):
    beta = float(g['meta/beta'])
    m = build_model(str(g['meta/kind']), bool(g['meta/perturb']), None if beta < 0 else beta, device)
    s = float(sum(p.detach().double().sum() for p in m.parameters()))
    assert abs(s - float(g['meta/param_sum'])) < 1e-6 * max(1.0, abs(s)), (s, float(g['meta/param_sum']))
    return m


def state_dict_cpu(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def conf_of(kind):
    return C.dtu_model_conf() if kind == 'dtu' else C.bmvs_model_conf()


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())
