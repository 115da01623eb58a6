"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (`/root/reference/volsdf/model/*`) on CPU.

Used by `oracle/make_golden.py` (to generate tests/golden/*.npz in the build container) and by the
`-m "not gpu"` tests that validate the oracle restatement when `/root/reference` is present.  Nothing
here runs on the GPU box: `/root/reference` does not exist there.

Shims: empty `imageio`/`skimage`/`omegaconf`/`GPUtil` modules (imported but
unused by the hot path: volsdf/utils/rend_util.py:2-3, helpers/help.py:8-10), and `.cuda()` made an
identity when no GPU is present (the reference hard-codes `.cuda()`, e.g. volsdf/model/network.py:198).
No reference file is modified or copied.
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get('SVS_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'volsdf', 'model'))


_loaded = {}


def load():
    """Returns a namespace with the reference's network / network_bg / ray_sampler / density / loss modules."""
    if _loaded:
        return _loaded['ns']
    if not available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    for name in ('imageio', 'skimage', 'GPUtil'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if 'omegaconf' not in sys.modules:
        m = types.ModuleType('omegaconf')
        m.OmegaConf = type('OmegaConf', (), {})
        sys.modules['omegaconf'] = m
    if 'nvidia_smi' not in sys.modules:
        sys.modules['nvidia_smi'] = types.ModuleType('nvidia_smi')
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    warnings.filterwarnings('ignore', message='.*weight_norm.*')
    from volsdf.model import network, network_bg, ray_sampler, density, embedder
    from volsdf.utils import rend_util
    try:
        from volsdf.model import loss
    except Exception:  # loss.py drags helpers.help; not needed for the hot path
        loss = None
    ns = types.SimpleNamespace(network=network, network_bg=network_bg, ray_sampler=ray_sampler,
                               density=density, embedder=embedder, rend_util=rend_util, loss=loss)
    _loaded['ns'] = ns
    return ns
