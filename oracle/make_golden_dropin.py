"""TEST INFRASTRUCTURE ONLY — goldens for the drop-in seam, generated from the UNMODIFIED reference on CPU:

  tests/golden/state_dict_keys.json   key -> shape of `VolSDFNetwork(conf).state_dict()` / `VolSDFNetworkBG(conf)...`
                                      (what a checkpoint written by volsdf/vsdf.py:189-191 contains)
  tests/golden/dtu_white_bkgd_r32.npz eval forward of the reference with `white_bkgd: true` (network.py:196-200,244-247)

Run in the build container (where /root/reference exists):  python -m oracle.make_golden_dropin
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
import svolsdf_b200.conf as C  # noqa: E402
import svolsdf_b200.scene as S  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def main():
    ns = ref_import.load()
    torch.set_num_threads(8)
    keys = {}
    torch.manual_seed(0)
    m = ns.network.VolSDFNetwork(C.dtu_model_conf())
    keys['dtu'] = {k: list(v.shape) for k, v in m.state_dict().items()}
    torch.manual_seed(0)
    mb = ns.network_bg.VolSDFNetworkBG(C.bmvs_model_conf())
    keys['bmvs'] = {k: list(v.shape) for k, v in mb.state_dict().items()}
    json.dump(keys, open(os.path.join(OUT, 'state_dict_keys.json'), 'w'), indent=0, sort_keys=True)

    R = 32
    torch.manual_seed(0)
    mw = ns.network.VolSDFNetwork(C.dtu_model_conf(white_bkgd=True, bg_color=(1.0, 0.5, 0.25)))
    S.perturb_(mw, seed=7, w_std=S.PERTURB_W, b_std=S.PERTURB_B, beta=0.05)
    mw.eval()
    inp = S.make_input('dtu', R)
    torch.manual_seed(123)
    out = mw(inp)      # (get_outputs needs autograd for the normals, also in eval)
    rec = {'out/' + k: out[k].detach().numpy() for k in ('rgb_values', 'depth_values', 'normal_map', 'weights')}
    rec['meta/param_sum'] = np.float64(sum(p.detach().double().sum() for p in mw.parameters()))
    np.savez_compressed(os.path.join(OUT, 'dtu_white_bkgd_r32.npz'), **rec)
    print('wrote', len(keys['dtu']), len(keys['bmvs']), 'keys;  white-bkgd acc range',
          float(out['weights'].sum(1).min()), float(out['weights'].sum(1).max()))


if __name__ == '__main__':
    main()
