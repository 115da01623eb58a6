"""Duck-typed stand-in for pyhocon's ConfigTree plus the two model configurations.

The reference builds its model from a pyhocon ConfigTree (volsdf/vsdf.py:25-26,92-93) and reads it
with get_int/get_float/get_bool/get_list/get_string/get_config (volsdf/model/network.py:195-204,
volsdf/model/network_bg.py:21-35).  pyhocon is not installed in this image, so tests and bench.py
use this class; a real ConfigTree works just as well with the model classes.

Values below restate config/vol/dtu.yaml:26-57 and config/vol/bmvs.yaml:26-76 of the reference.
"""
import copy

_MISSING = object()


class ConfTree(dict):
    """Nested dict with pyhocon-style typed getters and dotted keys."""

    def _lookup(self, key, default=_MISSING):
        node = self
        for part in key.split('.'):
            if isinstance(node, dict) and part in node:
                node = node[part]
            else:
                if default is _MISSING:
                    raise KeyError(key)
                return default
        return node

    def get_int(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else int(v)

    def get_float(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else float(v)

    def get_bool(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else bool(v)

    def get_string(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else str(v)

    def get_list(self, key, default=_MISSING):
        v = self._lookup(key, default)
        return v if v is None else list(v)

    def get_config(self, key, default=_MISSING):
        v = self._lookup(key, default)
        if isinstance(v, dict) and not isinstance(v, ConfTree):
            v = ConfTree(v)
        return v

    def get(self, key, default=None):
        return self._lookup(key, default)


def _wrap(d):
    out = ConfTree()
    for k, v in d.items():
        out[k] = _wrap(v) if isinstance(v, dict) else v
    return out


_DTU_MODEL = {
    'feature_vector_size': 256,
    'scene_bounding_sphere': 3.0,
    'implicit_network': {
        'd_in': 3, 'd_out': 1, 'dims': [256] * 8, 'geometric_init': True, 'bias': 0.6,
        'skip_in': [4], 'weight_norm': True, 'multires': 6, 'sphere_scale': 20.0,
    },
    'rendering_network': {
        'mode': 'idr', 'd_in': 9, 'd_out': 3, 'dims': [256] * 4, 'weight_norm': True,
        'multires_view': 1,
    },
    'density': {'params_init': {'beta': 0.1}, 'beta_min': 0.0001},
    'ray_sampler': {
        'near': 0.0, 'N_samples': 64, 'N_samples_eval': 128, 'N_samples_extra': 32,
        'eps': 0.1, 'beta_iters': 10, 'max_total_iters': 5,
    },
}

_BMVS_MODEL = {
    'feature_vector_size': 256,
    'scene_bounding_sphere': 3.0,
    'implicit_network': {
        'd_in': 3, 'd_out': 1, 'dims': [256] * 8, 'geometric_init': True, 'bias': 0.6,
        'skip_in': [4], 'weight_norm': True, 'multires': 6,
    },
    'rendering_network': {
        'mode': 'idr', 'd_in': 9, 'd_out': 3, 'dims': [256] * 4, 'weight_norm': True,
        'multires_view': 1,
    },
    'density': {'params_init': {'beta': 0.1}, 'beta_min': 0.0001},
    'ray_sampler': {
        'near': 0.0, 'N_samples': 64, 'N_samples_eval': 128, 'N_samples_extra': 32,
        'eps': 0.1, 'beta_iters': 10, 'max_total_iters': 5,
        'N_samples_inverse_sphere': 32, 'add_tiny': 1.0e-6,
    },
    'bg_network': {
        'feature_vector_size': 256,
        'implicit_network': {
            'd_in': 4, 'd_out': 1, 'dims': [256] * 8, 'geometric_init': False, 'bias': 0.0,
            'skip_in': [4], 'weight_norm': False, 'multires': 10,
        },
        'rendering_network': {
            'mode': 'nerf', 'd_in': 3, 'd_out': 3, 'dims': [128], 'weight_norm': False,
            'multires_view': 4,
        },
    },
}


def dtu_model_conf(white_bkgd=False, bg_color=(1.0, 1.0, 1.0)):
    """`model:` block of config/vol/dtu.yaml; `white_bkgd` / `bg_color` are the optional keys network.py:196-198 reads."""
    d = copy.deepcopy(_DTU_MODEL)
    if white_bkgd:
        d['white_bkgd'] = True
        d['bg_color'] = list(bg_color)
    return _wrap(d)


def bmvs_model_conf():
    """`model:` block of config/vol/bmvs.yaml."""
    return _wrap(copy.deepcopy(_BMVS_MODEL))
