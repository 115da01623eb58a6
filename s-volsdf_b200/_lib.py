"""ctypes binding of libsvolsdf_b200.so (C ABI: include/svs.h).

The library is plain `extern "C"` over raw device pointers — no torch types cross the boundary.  torch
is used here only to obtain `data_ptr()`s and the current CUDA stream.  There is NO fallback: if the
shared library is missing the import of any compute entry point raises, and every call checks the
returned status and raises `SvsError` with `svs_last_error()`.
"""
import ctypes as C
import os

import torch

SVS_MAX_LAYERS = 12
ENGINE_FP32, ENGINE_TC, ENGINE_TC_SPLIT = 0, 1, 2
NET_SDF, NET_RENDER = 0, 1
RENDER_IDR, RENDER_NERF = 0, 1
COMP_ABS_DENSITY, COMP_REVERSED, COMP_ZMAX_TAIL, COMP_FAST = 1, 2, 4, 8

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SVS_LIB_PATH') or os.path.join(_HERE, 'libsvolsdf_b200.so')


class SvsError(RuntimeError):
    pass


class MlpDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('n_layers', C.c_int32), ('d_in', C.c_int32), ('n_freqs', C.c_int32),
                ('skip_layer', C.c_int32), ('render_mode', C.c_int32), ('weight_norm', C.c_int32),
                ('in_dim', C.c_int32 * SVS_MAX_LAYERS), ('out_dim', C.c_int32 * SVS_MAX_LAYERS),
                ('sphere_radius', C.c_float), ('sphere_scale', C.c_float)]


class MlpParams(C.Structure):
    _fields_ = [('g', C.c_void_p * SVS_MAX_LAYERS), ('v', C.c_void_p * SVS_MAX_LAYERS),
                ('b', C.c_void_p * SVS_MAX_LAYERS)]


class SamplerCfg(C.Structure):
    _fields_ = [('near', C.c_float), ('far', C.c_float), ('eps', C.c_float), ('add_tiny', C.c_float),
                ('inv4logeps', C.c_float), ('beta_iters', C.c_int32), ('exact', C.c_int32)]


class MvsView(C.Structure):
    """svs_mvs_view (include/svs.h): one source view of VolOpt.cost_mapping"""
    _fields_ = [('cost', C.c_void_p), ('z_near', C.c_void_p), ('z_far', C.c_void_p),
                ('Dz', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('fx', C.c_float), ('fy', C.c_float), ('cx', C.c_float), ('cy', C.c_float), ('sk', C.c_float),
                ('c2w', C.c_float * 12), ('same_view', C.c_int32), ('view_id', C.c_int32)]


_P, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int32, C.c_float
_DESC, _PAR, _CFG = C.POINTER(MlpDesc), C.POINTER(MlpParams), C.POINTER(SamplerCfg)

# name -> (restype, argtypes); mirrors include/svs.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    'svs_last_error': (C.c_char_p, []),
    'svs_abi_version': (C.c_int, []),
    'svs_has_engine': (C.c_int, [C.c_int]),
    'svs_launch_count': (_I64, []),
    'svs_prof_enable': (C.c_int, [C.c_int]),
    'svs_prof_collect': (_I64, [C.c_char_p, _I64]),
    'svs_mlp_wbuf_floats': (_I64, [_DESC, C.c_int]),
    'svs_mlp_prepare': (C.c_int, [_DESC, _PAR, _P, C.c_int, _P]),
    'svs_mlp_param_grads': (C.c_int, [_DESC, _PAR, _P, _P, _PAR, _P]),
    'svs_sdf_ldy': (_I32, [_DESC]),
    'svs_sdf_ws_floats': (_I64, [_DESC, _I64, C.c_int, C.c_int]),
    'svs_sdf_saved_floats': (_I64, [_DESC, _I64, C.c_int]),
    'svs_sdf_bwd_ws_floats': (_I64, [_DESC, _I64, C.c_int]),
    'svs_sdf_forward': (C.c_int, [_DESC, _P, _P, _I64, _P, _P, _P, C.c_int, _P]),
    'svs_sdf_outputs_forward': (C.c_int, [_DESC, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, C.c_int, _P]),
    'svs_sdf_outputs_backward': (C.c_int, [_DESC, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    'svs_embed': (C.c_int, [_P, _I64, _I32, _I32, _P, _P]),
    'svs_render_saved_floats': (_I64, [_DESC, _I64, C.c_int]),
    'svs_render_ws_floats': (_I64, [_DESC, _I64, C.c_int]),
    'svs_render_forward': (C.c_int, [_DESC, _P, _P, _P, _P, _P, _I32, _I64, _P, _P, C.c_int, _P]),
    'svs_render_backward': (C.c_int, [_DESC, _P, _I64, _P, _P, _P, _P, _P, _I32, _P, _P, C.c_int, _P]),
    'svs_raygen': (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P]),
    'svs_sphere_intersections': (C.c_int, [_P, _P, _I64, _F, _P, _P, _P]),
    'svs_ray_points': (C.c_int, [_P, _P, _P, _I64, _I32, _I32, _P, _P]),
    'svs_depth2pts_outside': (C.c_int, [_P, _P, _P, _I64, _I32, _F, _P, _P, _P]),
    'svs_sampler_init': (C.c_int, [_CFG, _I64, _I32, _P, _P, _P, _P, _P, _P]),
    'svs_sampler_bound': (C.c_int, [_CFG, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P]),
    'svs_sampler_resample': (C.c_int, [_CFG, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P]),
    'svs_sampler_finalize': (C.c_int, [_CFG, _I64, _I32, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P]),
    'svs_composite_forward': (C.c_int, [_P, _P, _P, _P, _P, _F, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    'svs_composite_backward': (C.c_int, [_P, _P, _P, _P, _F, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    'svs_density_forward': (C.c_int, [_P, _I64, _I32, _P, _F, _P, _I32, _P, _P]),
    'svs_adam_step': (C.c_int, [_I32, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _I32, _P, _P, _P]),
    'svs_adam_step_allreduce': (C.c_int, [_I32, _P, _P, _P, _P, _I32, _I32, _P, _P, _I64, _P, _F, _F, _F, _F, _F, _I32, _P, _P, _P, _P]),
    'svs_density_backward': (C.c_int, [_P, _I64, _I32, _P, _F, _P, _I32, _P, _P, _P, _P]),
    'svs_cost_mapping': (C.c_int, [_P, _I64, _I32, C.POINTER(MvsView), _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    'svs_mvs_loss': (C.c_int, [_P, _I64, _I32, C.POINTER(MvsView), _I32, _I32, _I32, _I32, _P, _P, C.c_float, C.c_float, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvsError('%s not found — build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       '(or s-volsdf_b200/csrc/build.sh); svolsdf_b200 has no CPU/PyTorch fallback' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL); the tensor must be contiguous fp32/int on CUDA."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SvsError('svolsdf_b200 kernels need CUDA tensors (got %s); there is no CPU path' % t.device)
    if not t.is_contiguous():
        raise SvsError('non-contiguous tensor passed to a kernel')
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(rc, what):
    if rc != 0:
        msg = load().svs_last_error()
        raise SvsError('%s failed (%d): %s' % (what, rc, msg.decode() if msg else ''))


def call(name, *args):
    check(getattr(load(), name)(*args), name)


def make_desc(kind, in_dims, out_dims, d_in=3, n_freqs=0, skip_layer=-1, render_mode=0, weight_norm=True,
              sphere_radius=0.0, sphere_scale=1.0):
    d = MlpDesc()
    d.kind, d.n_layers, d.d_in, d.n_freqs = kind, len(in_dims), d_in, n_freqs
    d.skip_layer, d.render_mode, d.weight_norm = skip_layer, render_mode, 1 if weight_norm else 0
    for l, (i, o) in enumerate(zip(in_dims, out_dims)):
        d.in_dim[l], d.out_dim[l] = i, o
    d.sphere_radius, d.sphere_scale = sphere_radius, sphere_scale
    return d


def launch_count():
    return int(load().svs_launch_count())


def prof_enable(on):
    load().svs_prof_enable(1 if on else 0)


def prof_collect():
    """-> {kernel name: {'launches', 'ms', 'flops', 'bytes'}} for everything recorded since the last call."""
    buf = C.create_string_buffer(1 << 16)
    n = load().svs_prof_collect(buf, len(buf))
    if n < 0:
        raise SvsError('svs_prof_collect failed: %s' % load().svs_last_error().decode())
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms, flops, nbytes = line.split('\t')
        out[name] = {'launches': int(cnt), 'ms': float(ms), 'flops': float(flops), 'bytes': float(nbytes)}
    return out


def make_params(gs, vs, bs):
    p = MlpParams()
    for l, (g, v, b) in enumerate(zip(gs, vs, bs)):
        p.g[l] = ptr(g) if g is not None else None
        p.v[l] = ptr(v)
        p.b[l] = ptr(b)
    return p
