"""Density modules with the reference's interface (volsdf/model/density.py).

Inside VolSDFNetwork the density is fused into the compositor kernel; these modules exist so that
`model.density(sdf, beta=...)` and `model.density.get_beta()` keep working (vsdf.py:228-229, sampler API).
"""
import torch
import torch.nn as nn

from .. import _lib as L


class _DensityFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, beta_param, beta_min, beta_rows, abs_density):
        shape = sdf.shape
        s2 = sdf.detach().reshape(-1, shape[-1]).contiguous().float() if sdf.dim() > 1 else sdf.detach().reshape(1, -1).contiguous().float()
        R, S = s2.shape
        out = torch.empty_like(s2)
        bp = beta_param.detach().reshape(1).contiguous() if beta_param is not None else None
        br = beta_rows.detach().reshape(-1).contiguous().float() if beta_rows is not None else None
        if br is not None and br.numel() != R:
            raise L.SvsError('per-ray beta must have one value per row')
        L.call('svs_density_forward', L.ptr(s2), R, S, L.ptr(bp), float(beta_min), L.ptr(br), 1 if abs_density else 0,
               L.ptr(out), L.stream())
        ctx.save_for_backward(s2, bp if bp is not None else torch.empty(0), br if br is not None else torch.empty(0))
        ctx.meta = (shape, float(beta_min), abs_density, bp is not None, br is not None,
                    beta_param.shape if beta_param is not None else None)
        return out.reshape(shape)

    @staticmethod
    def backward(ctx, d_out):
        s2, bp, br = ctx.saved_tensors
        shape, beta_min, abs_density, has_bp, has_br, bshape = ctx.meta
        bp = bp if has_bp else None
        br = br if has_br else None
        R, S = s2.shape
        d_sdf = torch.empty_like(s2)
        d_beta = torch.zeros(1, device=s2.device) if (has_bp and not has_br and not abs_density) else None
        L.call('svs_density_backward', L.ptr(s2), R, S, L.ptr(bp), beta_min, L.ptr(br), 1 if abs_density else 0,
               L.ptr(d_out.reshape(R, S).contiguous()), L.ptr(d_sdf), L.ptr(d_beta), L.stream())
        return d_sdf.reshape(shape), (d_beta.reshape(bshape) if d_beta is not None else None), None, None, None


class Density(nn.Module):
    def __init__(self, params_init={}):
        super().__init__()
        for p in params_init:
            setattr(self, p, nn.Parameter(torch.tensor(params_init[p])))

    def forward(self, sdf, beta=None):
        return self.density_func(sdf, beta=beta)


class LaplaceDensity(Density):
    """alpha * Laplace(0, beta).cdf(-sdf), alpha = 1/beta (density.py:16-30)."""

    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__(params_init=params_init)
        self.beta_min = torch.tensor(beta_min)   # plain tensor, not a buffer, like the reference (density.py:19)

    def density_func(self, sdf, beta=None):
        if beta is None:
            return _DensityFn.apply(sdf, self.beta, float(self.beta_min), None, False)
        if beta.numel() == 1:   # explicit scalar beta (already includes beta_min)
            rows = beta.detach().reshape(1).expand(sdf.reshape(-1, sdf.shape[-1]).shape[0])
            return _DensityFn.apply(sdf, None, 0.0, rows, False)
        return _DensityFn.apply(sdf, None, 0.0, beta, False)

    def get_beta(self):
        return self.beta.abs() + self.beta_min.to(self.beta.device)


class AbsDensity(Density):
    """sigma = |sdf| (density.py:33-35), used for the inverted-sphere background."""

    def density_func(self, sdf, beta=None):
        return _DensityFn.apply(sdf, None, 0.0, None, True)
