"""The drop-in seam of SURVEY.md 8(b) without a GPU: the reference instantiates its model through a string-named class
(`utils.get_class(conf['train.model_class'])(conf=...)`, volsdf/vsdf.py:92-93, utils/general.py:10-16) and saves /
loads `model.state_dict()` (vsdf.py:189-191).  Construction and state-dict handling run on the CPU (only forward needs
the CUDA library)."""
import json
import os

import pytest
import torch

from helpers import GOLDEN
from oracle import ref_import
import svolsdf_b200.conf as C

KEYS = json.load(open(os.path.join(GOLDEN, 'state_dict_keys.json')))
needs_ref = pytest.mark.skipif(not ref_import.available(), reason='/root/reference is not mounted')


@pytest.mark.parametrize('kind', ['dtu', 'bmvs'])
def test_state_dict_keys_and_shapes_match_the_reference_classes(kind):
    """tests/golden/state_dict_keys.json was written from the reference's own classes (oracle/make_golden_dropin.py)"""
    from svolsdf_b200.model.network import VolSDFNetwork
    from svolsdf_b200.model.network_bg import VolSDFNetworkBG
    torch.manual_seed(0)
    m = VolSDFNetwork(C.dtu_model_conf()) if kind == 'dtu' else VolSDFNetworkBG(C.bmvs_model_conf())
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert ours == KEYS[kind]
    assert list(m.state_dict().keys()) == list(KEYS[kind].keys()) or sorted(ours) == sorted(KEYS[kind])
    assert sum(p.numel() for p in m.parameters()) == (797883 if kind == 'dtu' else 1361387)


@needs_ref
@pytest.mark.parametrize('kind,dotted', [('dtu', 'svolsdf_b200.model.network.VolSDFNetwork'),
                                         ('bmvs', 'svolsdf_b200.model.network_bg.VolSDFNetworkBG')])
def test_reference_get_class_instantiates_our_model_and_loads_a_reference_checkpoint(kind, dotted, tmp_path):
    ns = ref_import.load()
    from volsdf.utils import general          # the reference's own helper, unmodified
    conf = C.dtu_model_conf() if kind == 'dtu' else C.bmvs_model_conf()
    cls = general.get_class(dotted)
    torch.manual_seed(0)
    ours = cls(conf=conf)
    assert type(ours).__name__ == dotted.split('.')[-1] and type(ours).__module__.startswith('svolsdf_b200')
    # a checkpoint exactly as vsdf.py:189-191 writes it, from the reference's class with different weights
    torch.manual_seed(5)
    ref = (ns.network.VolSDFNetwork if kind == 'dtu' else ns.network_bg.VolSDFNetworkBG)(conf)
    path = os.path.join(str(tmp_path), 'latest.pth')
    torch.save({'epoch': 3, 'model_state_dict': ref.state_dict(), 'iter_step': 1234}, path)
    ck = torch.load(path)
    missing = ours.load_state_dict(ck['model_state_dict'], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for (ka, va), (kb, vb) in zip(sorted(ours.state_dict().items()), sorted(ref.state_dict().items())):
        assert ka == kb and torch.equal(va, vb), ka
    # same seed -> same initial weights as the reference's constructor (geometric init draws in the same RNG order)
    torch.manual_seed(0)
    a = cls(conf=conf)
    torch.manual_seed(0)
    b = (ns.network.VolSDFNetwork if kind == 'dtu' else ns.network_bg.VolSDFNetworkBG)(conf)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    # the optimizer of the loop (vsdf.py:102) takes the parameters as ordinary leaves
    opt = torch.optim.Adam(ours.parameters(), lr=5e-4)
    assert sum(p.numel() for g in opt.param_groups for p in g['params']) == (797883 if kind == 'dtu' else 1361387)


@needs_ref
def test_oracle_white_background_matches_the_reference():
    """the `white_bkgd` branch (network.py:196-200,244-247) of the oracle against the reference's recorded output"""
    import numpy as np
    from helpers import load_golden
    from oracle import volsdf_oracle as O
    import svolsdf_b200.scene as S
    from svolsdf_b200.model.network import VolSDFNetwork
    g = load_golden('dtu_white_bkgd_r32')
    conf = C.dtu_model_conf(white_bkgd=True, bg_color=(1.0, 0.5, 0.25))
    torch.manual_seed(0)
    m = VolSDFNetwork(conf)
    S.perturb_(m, seed=7, w_std=S.PERTURB_W, b_std=S.PERTURB_B, beta=0.05)
    assert abs(float(sum(p.detach().double().sum() for p in m.parameters())) - float(g['meta/param_sum'])) < 1e-6
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    torch.manual_seed(123)
    o = O.volsdf_forward(sd, conf, S.make_input('dtu', 32), False)
    assert np.abs(o['rgb_values'].detach().numpy() - g['out/rgb_values']).max() < 2e-5
    assert np.abs(o['weights'].detach().numpy() - g['out/weights']).max() < 2e-5
