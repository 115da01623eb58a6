"""torch.autograd.Functions over the C ABI (include/svs.h).

Each Function is one node of the reference's autograd graph replaced by hand-written forward/backward
kernels: SDF network with analytic input gradient and double backward (volsdf/model/network.py:71-123),
rendering network (:170-190), Laplace density + volume rendering + compositing (:281-295,239-276).
PyTorch only owns the memory and the graph bookkeeping; no tensor math of the hot path runs in ATen.
"""
import torch

from . import _lib as L
from ._lib import ptr


def _f32(*shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


def _contig(t):
    return None if t is None else t.contiguous()


# Bumped by optimisers that update parameters through raw pointers (svolsdf_b200.optim.FusedAdam): torch's per-tensor
# version counters do not see those writes, and NetHandle.prepare() caches the packed weights per (epoch, versions).
_WEIGHTS_EPOCH = [0]


def weights_changed():
    _WEIGHTS_EPOCH[0] += 1


class NetHandle(object):
    """Descriptor + parameter plumbing of one MLP (built once per nn.Module)."""

    def __init__(self, desc, layers, engine=L.ENGINE_FP32):
        self.desc = desc
        self.layers = layers            # list of (g or None, v, b) nn.Parameters, layer order
        self.engine = engine
        self.wbuf_floats = int(L.load().svs_mlp_wbuf_floats(desc, engine))
        if self.wbuf_floats < 0:
            raise L.SvsError('bad MLP descriptor: %s' % L.load().svs_last_error().decode())
        # the gradient accumulator shares the fp32 part of the layout (effective weights + biases); the packed fp16 images
        # behind it exist only in wbuf
        self.dwbuf_floats = int(L.load().svs_mlp_wbuf_floats(desc, L.ENGINE_FP32))
        self.ldy = int(L.load().svs_sdf_ldy(desc))
        self._cache_key, self._cache_wbuf = None, None

    def flat_params(self):
        out = []
        for g, v, b in self.layers:
            if g is not None:
                out.append(g)
            out.append(v)
            out.append(b)
        return out

    def params_struct(self, tensors=None):
        """tensors: optional replacement list in flat_params() order (used for gradient buffers)."""
        gs, vs, bs = [], [], []
        it = iter(tensors) if tensors is not None else None
        for g, v, b in self.layers:
            if it is None:
                gs.append(g.detach() if g is not None else None)
                vs.append(v.detach())
                bs.append(b.detach())
            else:
                gs.append(next(it) if g is not None else None)
                vs.append(next(it))
                bs.append(next(it))
        for t in vs + bs + [g for g in gs if g is not None]:
            if not t.is_contiguous():
                raise L.SvsError('MLP parameters must be contiguous')
        return L.make_params(gs, vs, bs)

    def prepare(self, device):
        """Effective weights (weight-norm applied) + packed operand images.  The reference recomputes W = g v / |v| in
        a pre-forward hook on EVERY call (network.py:64-65); here the packed buffer is reused until a parameter changes
        (one pack per optimiser step instead of one per entry point: sampler pass, main pass, rendering net ...).  A new
        buffer is allocated per pack, so buffers saved for a backward are never overwritten."""
        key = (_WEIGHTS_EPOCH[0], str(device), torch.cuda.is_current_stream_capturing() if torch.cuda.is_available() else False,
               tuple((p.data_ptr(), p._version) for p in self.flat_params()))
        if key == self._cache_key:
            return self._cache_wbuf
        wbuf = _f32(self.wbuf_floats, device=device)
        ps = self.params_struct()
        L.call('svs_mlp_prepare', self.desc, ps, ptr(wbuf), self.engine, L.stream())
        self._cache_key, self._cache_wbuf = key, wbuf
        return wbuf

    def param_grads(self, wbuf, dwbuf):
        grads = [torch.empty_like(p) for p in self.flat_params()]
        L.call('svs_mlp_param_grads', self.desc, self.params_struct(), ptr(wbuf), ptr(dwbuf),
               self.params_struct(grads), L.stream())
        return grads

    def needs_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.flat_params())


def sdf_forward_nograd(net, x, want_y, want_sdf):
    """ImplicitNetwork.forward / get_sdf_vals without autograd.  Returns (y (P,ldy) | None, sdf (P,1) | None)."""
    x = x.detach().contiguous().float()
    P = x.shape[0]
    dev = x.device
    wbuf = net.prepare(dev)
    ws = _f32(max(1, int(L.load().svs_sdf_ws_floats(net.desc, P, 0, net.engine))), device=dev)
    y = _f32(P, net.ldy, device=dev) if want_y else None
    sdf = _f32(P, 1, device=dev) if want_sdf else None
    if P == 0:
        return y, sdf
    L.call('svs_sdf_forward', net.desc, ptr(wbuf), ptr(x), P, ptr(y), ptr(sdf), ptr(ws), net.engine, L.stream())
    return y, sdf


class SdfOutputsFn(torch.autograd.Function):
    """x (P,d_in) -> y (P,ldy) raw outputs, sdf (P,1) clamped, grad (P,d_in) = d sdf / d x.
    `clamp`: True / False, or an int n: only the first n points take the bounding-sphere minimum (the ray samples of
    get_outputs(); the eikonal samples of gradient() ride behind them in the same launch).

    backward implements loss.backward() through get_outputs()/gradient(): first-order terms through y/sdf
    and the double-backward through the analytic gradient."""

    @staticmethod
    def forward(ctx, net, x, clamp, want_grad, train, *params):
        x = x.detach().contiguous().float()
        P = x.shape[0]
        dev = x.device
        wbuf = net.prepare(dev)
        lib = L.load()
        y = _f32(P, net.ldy, device=dev)
        sdf = _f32(P, 1, device=dev)
        grad = _f32(P, net.desc.d_in, device=dev) if want_grad else None
        saved = _f32(max(1, int(lib.svs_sdf_saved_floats(net.desc, P, net.engine))), device=dev) if train else None
        ws = _f32(max(1, int(lib.svs_sdf_ws_floats(net.desc, P, 0 if train else 1, net.engine))), device=dev)
        n_clamped = P if clamp is True else (0 if not clamp else int(clamp))   # True: all points, int: the leading n
        L.call('svs_sdf_outputs_forward', net.desc, ptr(wbuf), ptr(x), P, n_clamped, ptr(y), ptr(sdf),
               ptr(grad), ptr(saved), ptr(ws), net.engine, L.stream())
        ctx.net, ctx.clamp, ctx.P = net, n_clamped, P
        ctx.want_grad = bool(want_grad)
        if train:
            ctx.save_for_backward(x, y, saved, wbuf)
        if grad is None:
            grad = torch.zeros(P, net.desc.d_in, device=dev)
            ctx.mark_non_differentiable(grad)
        return y, sdf, grad

    @staticmethod
    def backward(ctx, dy, d_sdf, d_grad):
        net, P = ctx.net, ctx.P
        x, y, saved, wbuf = ctx.saved_tensors
        dev = x.device
        lib = L.load()
        dwbuf = torch.zeros(net.dwbuf_floats, dtype=torch.float32, device=dev)
        ws = _f32(max(1, int(lib.svs_sdf_bwd_ws_floats(net.desc, P, net.engine))), device=dev)
        dy, d_sdf, d_grad = _contig(dy), _contig(d_sdf), _contig(d_grad)
        if not ctx.want_grad:
            # autograd materialises a zero gradient for the (non-differentiable) dummy `grad` output; passing it on would
            # run the tangent sweep over the U tiles the forward never wrote (no reverse sweep without want_grad)
            d_grad = None
        L.call('svs_sdf_outputs_backward', net.desc, ptr(wbuf), ptr(x), P, ctx.clamp, ptr(saved), ptr(y),
               ptr(dy), ptr(d_sdf), ptr(d_grad), ptr(dwbuf), ptr(ws), net.engine, L.stream())
        grads = net.param_grads(wbuf, dwbuf)
        return (None, None, None, None, None) + tuple(grads)


def sdf_outputs(net, x, clamp=True, want_grad=True):
    return SdfOutputsFn.apply(net, x, clamp, want_grad, net.needs_grad(), *net.flat_params())


class RenderFn(torch.autograd.Function):
    """RenderingNetwork.forward.  `feat` may be a wider tensor (e.g. the SDF net's y) read at column
    `feat_col`; its gradient comes back with the same full width so no slicing copies are needed."""

    @staticmethod
    def forward(ctx, net, points, normals, view_dirs, feat, feat_col, train, *params):
        dev = view_dirs.device
        P = view_dirs.shape[0]
        view_dirs = view_dirs.detach().contiguous().float()
        idr = net.desc.render_mode == L.RENDER_IDR
        pts = points.detach().contiguous().float() if idr else None
        nrm = normals.detach().contiguous().float() if idr else None
        f = feat.detach()
        if f.stride(1) != 1:
            f = f.contiguous()
        ld_feat = f.stride(0)
        fptr = f.data_ptr() + 4 * feat_col
        lib = L.load()
        wbuf = net.prepare(dev)
        saved = _f32(max(1, int(lib.svs_render_saved_floats(net.desc, P, net.engine))), device=dev)
        rgb = _f32(P, net.desc.out_dim[net.desc.n_layers - 1], device=dev)
        L.call('svs_render_forward', net.desc, ptr(wbuf), ptr(pts), ptr(view_dirs), ptr(nrm), fptr, ld_feat, P,
               ptr(rgb), ptr(saved), net.engine, L.stream())
        ctx.net, ctx.P, ctx.feat_col, ctx.feat_shape, ctx.idr = net, P, feat_col, tuple(feat.shape), idr
        if train:
            ctx.save_for_backward(saved, rgb, wbuf)
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        net, P = ctx.net, ctx.P
        saved, rgb, wbuf = ctx.saved_tensors
        dev = rgb.device
        lib = L.load()
        dwbuf = torch.zeros(net.dwbuf_floats, dtype=torch.float32, device=dev)
        ws = _f32(max(1, int(lib.svs_render_ws_floats(net.desc, P, net.engine))), device=dev)
        d_normals = _f32(P, 3, device=dev) if ctx.idr else None
        # the kernel writes the F feature columns; only the columns around them (the sdf column / row padding of the SDF
        # net's y when the features are read in place) need zeros — not a fill of the whole (P, ldy) tensor
        d_feat = torch.empty(ctx.feat_shape, dtype=torch.float32, device=dev)
        F_ = net.desc.in_dim[0] - (3 * (1 + 2 * net.desc.n_freqs) + (6 if ctx.idr else 0))
        if ctx.feat_col > 0:
            d_feat[:, :ctx.feat_col].zero_()
        if ctx.feat_col + F_ < d_feat.shape[1]:
            d_feat[:, ctx.feat_col + F_:].zero_()
        if d_feat.shape[0] > P:      # rows behind the rendered points (the eikonal samples ride at the tail of y)
            d_feat[P:].zero_()
        L.call('svs_render_backward', net.desc, ptr(wbuf), P, ptr(saved), ptr(rgb), ptr(d_rgb.contiguous()),
               ptr(d_normals), d_feat.data_ptr() + 4 * ctx.feat_col, d_feat.stride(0), ptr(dwbuf), ptr(ws),
               net.engine, L.stream())
        grads = net.param_grads(wbuf, dwbuf)
        return (None, None, d_normals, None, d_feat, None, None) + tuple(grads)


def render(net, points, normals, view_dirs, feat, feat_col=0):
    train = net.needs_grad() or (torch.is_grad_enabled() and (
        feat.requires_grad or (normals is not None and normals.requires_grad)))
    return RenderFn.apply(net, points, normals, view_dirs, feat, feat_col, train, *net.flat_params())


class CompositeFn(torch.autograd.Function):
    """density -> transmittance -> weights -> rgb/depth(/normal) maps, one fused kernel each way."""

    @staticmethod
    def forward(ctx, z, sdf, rgb, beta_param, beta_min, depth_scale, normals, z_max, flags):
        dev = z.device
        R, S = z.shape
        z = z.detach().contiguous().float()
        sdf_c = sdf.detach().reshape(R, S).contiguous().float()
        rgb_c = rgb.detach().reshape(R, S, 3).contiguous().float() if rgb is not None else None
        ds = depth_scale.detach().reshape(R).contiguous().float() if depth_scale is not None else None
        nrm = normals.detach().reshape(R, S, 3).contiguous().float() if normals is not None else None
        zm = z_max.detach().reshape(R).contiguous().float() if z_max is not None else None
        bp = beta_param.detach().reshape(1).contiguous() if beta_param is not None else None
        weights = _f32(R, S, device=dev)
        rgb_values = _f32(R, 3, device=dev) if rgb_c is not None else None
        depth_values = _f32(R, 1, device=dev)
        normal_map = _f32(R, 3, device=dev) if nrm is not None else None
        bg_trans = _f32(R, device=dev) if (flags & L.COMP_ZMAX_TAIL) else None
        L.call('svs_composite_forward', ptr(z), ptr(sdf_c), ptr(rgb_c), ptr(nrm), ptr(bp), float(beta_min), ptr(ds),
               ptr(zm), R, S, flags, ptr(weights), ptr(rgb_values), ptr(depth_values), ptr(normal_map),
               ptr(bg_trans), L.stream())
        ctx.flags, ctx.beta_min, ctx.R, ctx.S = flags, float(beta_min), R, S
        ctx.has = (rgb_c is not None, ds is not None, zm is not None, bp is not None)
        ctx.shapes = (tuple(sdf.shape), tuple(rgb.shape) if rgb is not None else None,
                      tuple(beta_param.shape) if beta_param is not None else None)
        ctx.save_for_backward(*[t if t is not None else torch.empty(0, device=dev) for t in (z, sdf_c, rgb_c, bp, ds, zm)])
        outs = [weights,
                rgb_values if rgb_values is not None else torch.zeros(R, 3, device=dev),
                depth_values,
                normal_map if normal_map is not None else torch.zeros(R, 3, device=dev),
                bg_trans if bg_trans is not None else torch.zeros(R, device=dev)]
        ctx.mark_non_differentiable(outs[3])
        return tuple(outs)

    @staticmethod
    def backward(ctx, d_weights, d_rgb_values, d_depth_values, d_normal_map, d_bg_trans):
        z, sdf_c, rgb_c, bp, ds, zm = ctx.saved_tensors
        has_rgb, has_ds, has_zm, has_bp = ctx.has
        rgb_c = rgb_c if has_rgb else None
        ds = ds if has_ds else None
        zm = zm if has_zm else None
        bp = bp if has_bp else None
        dev = z.device
        R, S = ctx.R, ctx.S
        d_sdf = _f32(R, S, device=dev)
        d_rgb = _f32(R, S, 3, device=dev) if has_rgb else None
        d_beta = torch.zeros(1, dtype=torch.float32, device=dev) if has_bp else None
        drv = _contig(d_rgb_values) if has_rgb else None
        ddv = _contig(d_depth_values)
        dbt = _contig(d_bg_trans) if (ctx.flags & L.COMP_ZMAX_TAIL) else None
        L.call('svs_composite_backward', ptr(z), ptr(sdf_c), ptr(rgb_c), ptr(bp), ctx.beta_min, ptr(ds), ptr(zm), R, S,
               ctx.flags, ptr(drv), ptr(ddv), ptr(_contig(d_weights)), ptr(dbt), ptr(d_sdf), ptr(d_rgb), ptr(d_beta),
               L.stream())
        sdf_shape, rgb_shape, beta_shape = ctx.shapes
        return (None, d_sdf.reshape(sdf_shape), d_rgb.reshape(rgb_shape) if has_rgb else None,
                d_beta.reshape(beta_shape) if has_bp else None, None, None, None, None, None)


def composite(z, sdf, rgb, beta_param, beta_min, depth_scale=None, normals=None, z_max=None, flags=0):
    """Returns weights (R,S), rgb_values (R,3), depth_values (R,1), normal_map (R,3), bg_trans (R,)."""
    return CompositeFn.apply(z, sdf, rgb, beta_param, beta_min, depth_scale, normals, z_max, flags)


# ---- ray helpers (no autograd: ray geometry carries no gradient in the reference's loop) -------------

def raygen(uv, pose, intrinsics):
    """uv (R,2), pose (4,4), intrinsics (4,4) -> ray_dirs (R,3), cam_loc (R,3), depth_scale (R,1)."""
    uv = uv.detach().contiguous().float()
    R = uv.shape[0]
    dev = uv.device
    dirs, cam, ds = _f32(R, 3, device=dev), _f32(R, 3, device=dev), _f32(R, 1, device=dev)
    L.call('svs_raygen', ptr(uv), ptr(pose.detach().contiguous().float()), ptr(intrinsics.detach().contiguous().float()),
           R, ptr(dirs), ptr(cam), ptr(ds), L.stream())
    return dirs, cam, ds


def ray_points(cam_loc, ray_dirs, z):
    """(R,3),(R,3),(R,S) -> (R,S,3)"""
    z = z.detach()
    if z.stride(1) != 1:
        z = z.contiguous()
    R, S = z.shape
    pts = _f32(R, S, 3, device=z.device)
    L.call('svs_ray_points', ptr(cam_loc.contiguous()), ptr(ray_dirs.contiguous()), z.data_ptr(), R, S, z.stride(0),
           ptr(pts), L.stream())
    return pts


def sphere_intersections(cam_loc, ray_dirs, r):
    R = cam_loc.shape[0]
    dev = cam_loc.device
    nf = _f32(R, 2, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call('svs_sphere_intersections', ptr(cam_loc.contiguous()), ptr(ray_dirs.contiguous()), R, float(r), ptr(nf),
           ptr(bad), L.stream())
    return nf, bad


def depth2pts_outside(cam_loc, ray_dirs, depth, radius):
    """(R,3),(R,3),(R,S) inverse depths -> pts (R,S,4), depth_real (R,S)  (network_bg.py:182-214)"""
    depth = depth.detach().contiguous().float()
    R, S = depth.shape
    dev = depth.device
    pts, dr = _f32(R, S, 4, device=dev), _f32(R, S, device=dev)
    L.call('svs_depth2pts_outside', ptr(cam_loc.contiguous()), ptr(ray_dirs.contiguous()), ptr(depth), R, S,
           float(radius), ptr(pts), ptr(dr), L.stream())
    return pts, dr
