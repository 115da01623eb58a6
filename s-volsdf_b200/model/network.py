"""ImplicitNetwork / RenderingNetwork / VolSDFNetwork with the reference's interface
(volsdf/model/network.py), computed by the sm_100a kernels behind include/svs.h.

Drop-in contract: same constructor kwargs (the YAML keys of config/vol/*.yaml), same
`state_dict` keys (`implicit_network.lin{l}.{weight_g,weight_v,bias}`, `rendering_network.lin{l}...`,
`density.beta`), same `forward(input, fast=-1)` dict; parameters are ordinary leaf nn.Parameters, and
`rgb_values`, `depth_values`, `weights`, `grad_theta` are connected to autograd through hand-written
backward kernels (including the double backward through d sdf/dx).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .. import functional as F
from .density import LaplaceDensity
from .embedder import get_embedder
from .ray_sampler import ErrorBoundSampler, RefRng


def _layer_params(lin):
    if hasattr(lin, 'weight_g'):
        return (lin.weight_g, lin.weight_v, lin.bias)
    return (None, lin.weight, lin.bias)


class ImplicitNetwork(nn.Module):
    """PE + softplus(beta=100) MLP with a skip connection and geometric init (network.py:10-131)."""

    def __init__(self, feature_vector_size, sdf_bounding_sphere, d_in, d_out, dims, geometric_init=True, bias=1.0,
                 skip_in=(), weight_norm=True, multires=0, sphere_scale=1.0):
        super().__init__()
        self.sdf_bounding_sphere = sdf_bounding_sphere
        self.sphere_scale = sphere_scale
        self.d_in = d_in
        self.multires = multires
        self.feature_vector_size = feature_vector_size
        self.d_out = d_out
        widths = [d_in] + list(dims) + [d_out + feature_vector_size]
        self.embed_fn = None
        if multires > 0:
            self.embed_fn, widths[0] = get_embedder(multires, input_dims=d_in)
        self.num_layers = len(widths)
        self.skip_in = tuple(skip_in)
        self._in_dims, self._out_dims = [], []
        pe_extra = widths[0] - 3
        for l in range(self.num_layers - 1):
            n_out = widths[l + 1] - widths[0] if (l + 1) in self.skip_in else widths[l + 1]
            lin = nn.Linear(widths[l], n_out)
            if geometric_init:
                last = (l == self.num_layers - 2)
                if last:    # sphere of radius `bias`: mean sqrt(pi)/sqrt(fan_in), tiny spread
                    nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(widths[l]), std=0.0001)
                    nn.init.constant_(lin.bias, -bias)
                else:
                    nn.init.constant_(lin.bias, 0.0)
                    std = np.sqrt(2) / np.sqrt(n_out)
                    if multires > 0 and l == 0:          # only the raw xyz columns start non-zero
                        nn.init.constant_(lin.weight[:, 3:], 0.0)
                        nn.init.normal_(lin.weight[:, :3], 0.0, std)
                    elif multires > 0 and l in self.skip_in:  # re-injected PE columns start at zero
                        nn.init.normal_(lin.weight, 0.0, std)
                        nn.init.constant_(lin.weight[:, -pe_extra:], 0.0)
                    else:
                        nn.init.normal_(lin.weight, 0.0, std)
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, 'lin' + str(l), lin)
            self._in_dims.append(widths[l])
            self._out_dims.append(n_out)
        self.weight_norm = weight_norm
        self.softplus = nn.Softplus(beta=100)   # kept for interface parity; the kernels fuse it
        self.engine = L.ENGINE_FP32
        self._net = None
        self._net_raw = None
        if len(self.skip_in) > 1:
            raise NotImplementedError('one skip connection is supported (config/vol/*.yaml use skip_in=[4])')

    # -- kernel plumbing --------------------------------------------------------------------------
    def net(self):
        if self._net is None or self._net.engine != self.engine:
            desc = L.make_desc(L.NET_SDF, self._in_dims, self._out_dims, d_in=self.d_in, n_freqs=self.multires,
                               skip_layer=self.skip_in[0] if self.skip_in else -1, weight_norm=self.weight_norm,
                               sphere_radius=float(self.sdf_bounding_sphere), sphere_scale=float(self.sphere_scale))
            layers = [_layer_params(getattr(self, 'lin' + str(l))) for l in range(self.num_layers - 1)]
            self._net = F.NetHandle(desc, layers, self.engine)
        return self._net

    @property
    def n_out(self):
        return self._out_dims[-1]

    def outputs_fused(self, x, clamp=True):
        """-> y (P, ldy) raw outputs, sdf (P,1) clamped, gradients (P,d_in).  Internal fused entry."""
        return F.sdf_outputs(self.net(), x, clamp=clamp, want_grad=True)

    # -- reference interface ----------------------------------------------------------------------
    def forward(self, input):
        net = self.net()
        if net.needs_grad():
            y, _, _ = F.sdf_outputs(net, input, clamp=False, want_grad=False)
        else:
            y, _ = F.sdf_forward_nograd(net, input, True, False)
        return y[:, :self.n_out]

    def gradient(self, x, with_sdf=False):
        y, _, g = F.sdf_outputs(self.net(), x, clamp=False, want_grad=True)
        if with_sdf:
            return g, y[:, :1]
        return g

    def get_outputs(self, x):
        y, sdf, g = F.sdf_outputs(self.net(), x, clamp=True, want_grad=True)
        return sdf, y[:, 1:self.n_out], g

    def sdf_only(self, x):
        """forward(x)[:, :1] without computing the feature columns (no sphere clamp, no autograd): the bulk query of
        mesh extraction, `sdf = lambda x: model.implicit_network(x)[:, 0]` (eval_vsdf.py:115,132; utils/plots.py:61)."""
        if self._net_raw is None or self._net_raw.engine != self.engine:
            desc = L.make_desc(L.NET_SDF, self._in_dims, self._out_dims, d_in=self.d_in, n_freqs=self.multires,
                               skip_layer=self.skip_in[0] if self.skip_in else -1, weight_norm=self.weight_norm,
                               sphere_radius=0.0, sphere_scale=1.0)
            layers = [_layer_params(getattr(self, 'lin' + str(l))) for l in range(self.num_layers - 1)]
            self._net_raw = F.NetHandle(desc, layers, self.engine)
        with torch.no_grad():
            return F.sdf_forward_nograd(self._net_raw, x, False, True)[1]

    def get_sdf_vals(self, x):
        net = self.net()
        if net.needs_grad():
            return F.sdf_outputs(net, x, clamp=True, want_grad=False)[1]
        return F.sdf_forward_nograd(net, x, False, True)[1]


class RenderingNetwork(nn.Module):
    """ReLU MLP + sigmoid on cat[points, PE(view), normals, features] ('idr') or cat[PE(view), features]
    ('nerf') (network.py:134-190)."""

    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True, multires_view=0):
        super().__init__()
        self.mode = mode
        self.multires_view = multires_view
        self.feature_vector_size = feature_vector_size
        widths = [d_in + feature_vector_size] + list(dims) + [d_out]
        self.embedview_fn = None
        if multires_view > 0:
            self.embedview_fn, view_ch = get_embedder(multires_view)
            widths[0] += view_ch - 3
        self.num_layers = len(widths)
        for l in range(self.num_layers - 1):
            lin = nn.Linear(widths[l], widths[l + 1])
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, 'lin' + str(l), lin)
        self._in_dims, self._out_dims = widths[:-1], widths[1:]
        self.weight_norm = weight_norm
        self.relu = nn.ReLU()
        self.sigmoid = nn.Sigmoid()
        self.engine = L.ENGINE_FP32
        self._net = None
        if mode not in ('idr', 'nerf'):
            raise ValueError('unknown rendering mode %r' % (mode,))

    def net(self):
        if self._net is None or self._net.engine != self.engine:
            desc = L.make_desc(L.NET_RENDER, self._in_dims, self._out_dims, n_freqs=self.multires_view,
                               render_mode=L.RENDER_IDR if self.mode == 'idr' else L.RENDER_NERF,
                               weight_norm=self.weight_norm)
            layers = [_layer_params(getattr(self, 'lin' + str(l))) for l in range(self.num_layers - 1)]
            self._net = F.NetHandle(desc, layers, self.engine)
        return self._net

    def forward(self, points, normals, view_dirs, feature_vectors, _feat_col=0):
        return F.render(self.net(), points, normals, view_dirs, feature_vectors, _feat_col)


class VolSDFNetwork(nn.Module):
    """network.py:192-295"""

    def __init__(self, conf):
        super().__init__()
        self.feature_vector_size = conf.get_int('feature_vector_size')
        self.scene_bounding_sphere = conf.get_float('scene_bounding_sphere', default=1.0)
        self.white_bkgd = conf.get_bool('white_bkgd', default=False)
        self.bg_color = torch.tensor(conf.get_list('bg_color', default=[1.0, 1.0, 1.0])).float()
        self.implicit_network = ImplicitNetwork(self.feature_vector_size,
                                                0.0 if self.white_bkgd else self.scene_bounding_sphere,
                                                **conf.get_config('implicit_network'))
        self.rendering_network = RenderingNetwork(self.feature_vector_size, **conf.get_config('rendering_network'))
        self.density = LaplaceDensity(**conf.get_config('density'))
        self.ray_sampler = ErrorBoundSampler(self.scene_bounding_sphere, **conf.get_config('ray_sampler'))
        self.last_rng = None

    def set_engine(self, engine):
        self.implicit_network.engine = engine
        self.rendering_network.engine = engine
        return self

    def _comp_flags(self):
        """The tcgen05 engines composite with the bandwidth-bound arithmetic (MUFU exp, fp32 scans, lean kernels at 68 % /
        92 % of the HBM roofline: weights within 1e-6 of the canonical ones, far inside the 1e-3 contract); the fp32 parity
        engine keeps the canonical arithmetic (libm exp, fp64 prefix sums) that the bit-exactness tests pin."""
        return L.COMP_FAST if self.implicit_network.engine in (L.ENGINE_TC, L.ENGINE_TC_SPLIT) else 0

    def _rays(self, uv, pose, intrinsics):
        dirs, cams, scales = [], [], []
        for b in range(uv.shape[0]):
            d, c, s = F.raygen(uv[b], pose[b], intrinsics[b])
            dirs.append(d)
            cams.append(c)
            scales.append(s)
        if len(dirs) == 1:
            return dirs[0], cams[0], scales[0]
        # the reference takes depth_scale from batch element 0 only (network.py:217)
        return torch.cat(dirs, 0), torch.cat(cams, 0), scales[0]

    def forward(self, input, fast=-1):
        intrinsics, uv, pose = input['intrinsics'], input['uv'], input['pose']
        iter_step = input.get('iter_step', 1)
        if not self.training:
            with torch.no_grad():   # the reference's callers detach every eval output (vsdf.py:250-256)
                return self._forward(intrinsics, uv, pose, iter_step, fast)
        return self._forward(intrinsics, uv, pose, iter_step, fast)

    def _forward(self, intrinsics, uv, pose, iter_step, fast):
        dev = uv.device
        ray_dirs, cam_loc, depth_scale = self._rays(uv, pose, intrinsics)
        R = ray_dirs.shape[0]
        rng = self.rng_source if getattr(self, 'rng_source', None) is not None else RefRng(dev)
        self.last_rng = rng
        z_vals, z_samples_eik = self.ray_sampler.get_z_vals(ray_dirs, cam_loc, self, fast=fast, iter_step=iter_step,
                                                            _rng=rng)
        self.last_z = (z_vals, z_samples_eik)    # kept for parity tests (inject the same samples into the oracle)
        S = z_vals.shape[1]
        points = F.ray_points(cam_loc, ray_dirs, z_vals)               # (R,S,3)
        points_flat = points.reshape(-1, 3)
        dirs_flat = ray_dirs.unsqueeze(1).expand(R, S, 3).reshape(-1, 3)

        n_main = points_flat.shape[0]
        if self.training:
            # eikonal samples (network.py:258-268): R uniform points in the bounding box + one near-surface point per
            # ray.  They go through the SDF net in the SAME launches as the ray samples (unclamped tail of the batch).
            eikonal_points = rng.uniform((R, 3), -self.scene_bounding_sphere, self.scene_bounding_sphere)
            eik_near_points = F.ray_points(cam_loc, ray_dirs, z_samples_eik).reshape(-1, 3)
            all_points = torch.cat([points_flat, eikonal_points, eik_near_points], 0)
            y, sdf_all, grad_all = self.implicit_network.outputs_fused(all_points, clamp=n_main)
            sdf, gradients, grad_theta = sdf_all[:n_main], grad_all[:n_main], grad_all[n_main:]
        else:
            y, sdf, gradients = self.implicit_network.outputs_fused(points_flat, clamp=True)
        rgb_flat = self.rendering_network(points_flat, gradients, dirs_flat, y, _feat_col=1)
        weights, rgb_values, depth_values, normal_map, _ = F.composite(
            z_vals, sdf, rgb_flat, self.density.beta, float(self.density.beta_min), depth_scale,
            normals=None if self.training else gradients, flags=self._comp_flags())

        if self.white_bkgd:
            acc_map = torch.sum(weights, -1)
            rgb_values = rgb_values + (1. - acc_map[..., None]) * self.bg_color.to(dev).unsqueeze(0)

        output = {
            'rgb_values': rgb_values,
            'depth_values': depth_values,
            'depth_vals': z_vals * depth_scale,
            'weights': weights,
            'xyz': points,
        }
        if self.training:
            output['grad_theta'] = grad_theta
        else:
            output['normal_map'] = normal_map
        return output

    def volume_rendering(self, z_vals, sdf):
        """weights, dists (network.py:281-295)"""
        weights, _, _, _, _ = F.composite(z_vals, sdf, None, self.density.beta, float(self.density.beta_min))
        dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1],
                           torch.full((z_vals.shape[0], 1), 1e10, device=z_vals.device)], -1)
        return weights, dists
