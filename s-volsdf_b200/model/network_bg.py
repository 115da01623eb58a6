"""VolSDFNetworkBG with the reference's interface (volsdf/model/network_bg.py): foreground VolSDF inside
the bounding sphere + NeRF++ inverted-sphere background (config/vol/bmvs.yaml).

Foreground and background each run through the same kernels as VolSDFNetwork (SDF MLP, rendering MLP,
fused compositor with the z_max tail / AbsDensity variants); only the final blend of the two per-ray
results (a handful of (R,1)/(R,129) element-wise ops, network_bg.py:103-114) is left to torch autograd.
"""
import torch
import torch.nn as nn

from .. import _lib as L
from .. import functional as F
from .density import AbsDensity, LaplaceDensity
from .network import ImplicitNetwork, RenderingNetwork
from .ray_sampler import ErrorBoundSampler, RefRng


class VolSDFNetworkBG(nn.Module):
    def __init__(self, conf):
        super().__init__()
        self.feature_vector_size = conf.get_int('feature_vector_size')
        self.scene_bounding_sphere = conf.get_float('scene_bounding_sphere', default=1.0)
        # foreground object (no sphere clamp: the background takes over outside, network_bg.py:25)
        self.implicit_network = ImplicitNetwork(self.feature_vector_size, 0.0, **conf.get_config('implicit_network'))
        self.rendering_network = RenderingNetwork(self.feature_vector_size, **conf.get_config('rendering_network'))
        self.density = LaplaceDensity(**conf.get_config('density'))
        self.ray_sampler = ErrorBoundSampler(self.scene_bounding_sphere, inverse_sphere_bg=True,
                                             **conf.get_config('ray_sampler'))
        # background
        bg_fvs = conf.get_int('bg_network.feature_vector_size')
        self.bg_implicit_network = ImplicitNetwork(bg_fvs, 0.0, **conf.get_config('bg_network.implicit_network'))
        self.bg_rendering_network = RenderingNetwork(bg_fvs, **conf.get_config('bg_network.rendering_network'))
        self.bg_density = AbsDensity(**conf.get_config('bg_network.density', default={}))
        self.last_rng = None

    def set_engine(self, engine):
        for m in (self.implicit_network, self.rendering_network, self.bg_implicit_network, self.bg_rendering_network):
            m.engine = engine
        return self

    def _comp_flags(self):
        """see VolSDFNetwork._comp_flags: bandwidth-bound compositing arithmetic with the tcgen05 engine"""
        return L.COMP_FAST if self.implicit_network.engine in (L.ENGINE_TC, L.ENGINE_TC_SPLIT) else 0

    def forward(self, input, fast=-1):
        if not self.training:
            with torch.no_grad():
                return self._forward(input, fast)
        return self._forward(input, fast)

    def _forward(self, input, fast):
        intrinsics, uv, pose = input['intrinsics'], input['uv'], input['pose']
        dev = uv.device
        ray_dirs, cam_loc, depth_scale = F.raygen(uv[0], pose[0], intrinsics[0])
        R = ray_dirs.shape[0]
        rng = self.rng_source if getattr(self, 'rng_source', None) is not None else RefRng(dev)
        self.last_rng = rng
        (z_all, z_vals_bg), z_samples_eik = self.ray_sampler.get_z_vals(ray_dirs, cam_loc, self, fast=fast, _rng=rng)
        self.last_z = ((z_all, z_vals_bg), z_samples_eik)
        z_max = z_all[:, -1].contiguous()
        z_vals = z_all[:, :-1].contiguous()
        S = z_vals.shape[1]
        points = F.ray_points(cam_loc, ray_dirs, z_vals)
        points_flat = points.reshape(-1, 3)
        view = ray_dirs
        if not self.training:   # nearest training view direction (network_bg.py:70-74)
            view, _, _ = F.raygen(uv[0], input['near_pose'].to(dev)[0], intrinsics[0])
        dirs_flat = view.unsqueeze(1).expand(R, S, 3).reshape(-1, 3)

        y, sdf, gradients = self.implicit_network.outputs_fused(points_flat, clamp=False)
        rgb_flat = self.rendering_network(points_flat, gradients, dirs_flat, y, _feat_col=1)
        weights, fg_rgb_values, depth_values, normal_map, bg_transmittance = F.composite(
            z_vals, sdf, rgb_flat, self.density.beta, float(self.density.beta_min), depth_scale,
            normals=None if self.training else gradients, z_max=z_max, flags=L.COMP_ZMAX_TAIL | self._comp_flags())

        # background: inverted-sphere samples 1 -> 0 (network_bg.py:82-101)
        Sb = z_vals_bg.shape[1]
        z_bg = torch.flip(z_vals_bg, dims=[-1]).contiguous()
        bg_points, bg_depth_vals = F.depth2pts_outside(cam_loc, ray_dirs, z_bg, self.scene_bounding_sphere)
        bg_dirs_flat = view.unsqueeze(1).expand(R, Sb, 3).reshape(-1, 3)
        bnet = self.bg_implicit_network
        if bnet.net().needs_grad():
            yb, _, _ = F.sdf_outputs(bnet.net(), bg_points.reshape(-1, 4), clamp=False, want_grad=False)
        else:
            yb, _ = F.sdf_forward_nograd(bnet.net(), bg_points.reshape(-1, 4), True, False)
        bg_rgb_flat = self.bg_rendering_network(None, None, bg_dirs_flat, yb, _feat_col=1)
        bg_weights, bg_rgb_values, _, _, _ = F.composite(
            z_bg, yb[:, 0].reshape(R, Sb), bg_rgb_flat, None, 0.0, None,
            flags=L.COMP_ABS_DENSITY | L.COMP_REVERSED | self._comp_flags())

        # blend (network_bg.py:103-114)
        weights_all = torch.cat([weights, bg_transmittance[:, None] * bg_weights], 1)
        depth_vals_all = depth_scale * torch.cat([z_vals, bg_depth_vals], 1)
        depth_values_all = torch.sum(weights_all * depth_vals_all, 1, keepdim=True) / \
            (weights_all.sum(dim=1, keepdim=True) + 1e-8)
        rgb_values = fg_rgb_values + bg_transmittance.unsqueeze(-1) * bg_rgb_values

        output = {
            'rgb_values': rgb_values,
            'depth_values_all': depth_values_all,
            'depth_values': depth_values,
            'depth_vals': z_vals * depth_scale,
            'weights': weights,
            'xyz': points.detach(),
        }
        if self.training:
            eikonal_points = rng.uniform((R, 3), -self.scene_bounding_sphere, self.scene_bounding_sphere)
            eik_near_points = F.ray_points(cam_loc, ray_dirs, z_samples_eik).reshape(-1, 3)
            eikonal_points = torch.cat([eikonal_points, eik_near_points], 0)
            output['grad_theta'] = self.implicit_network.gradient(eikonal_points)
        else:
            output['normal_map'] = normal_map
        return output

    def depth2pts_outside(self, ray_o, ray_d, depth):
        """ray_o, ray_d (...,3), depth (...) -> pts (...,4), depth_real (...)  (network_bg.py:182-214)"""
        shp = depth.shape
        o = ray_o.reshape(-1, 3).contiguous()
        d = ray_d.reshape(-1, 3).contiguous()
        pts, dr = F.depth2pts_outside(o, d, depth.reshape(-1, 1), self.scene_bounding_sphere)
        return pts.reshape(tuple(shp) + (4,)), dr.reshape(shp)
