"""`VolSDFLoss` with the reference's constructor and result dictionary (volsdf/model/loss.py:15-115), so that
`config/vol/*.yaml`'s `loss_class` can point here next to `model_class`.

Terms: photometric loss (the class named by `rgb_loss`, mean reduction), eikonal term on `grad_theta`, the
generalised cross-entropy between the rendering weights and the MVS probabilities `pi * pj` that
`svolsdf_b200.mvs.CostMapper` looks up, and the annealed sparsity prior on the rendered depth.  Everything here is
O(rays x samples) glue on tensors the kernels produced; the gradients it sends back (`dL/drgb_values`,
`dL/dweights`, `dL/ddepth_values`, `dL/dgrad_theta`) are what `svs_composite_backward` and the MLP backward chains
consume.
"""
import importlib

import torch
from torch import nn


def _class_by_path(path):
    mod, _, name = path.rpartition('.')
    return getattr(importlib.import_module(mod), name)


class VolSDFLoss(nn.Module):
    def __init__(self, rgb_loss, eikonal_weight, rgb_weight=1., mvs_weight=0., sparse_weight=0., anneal_rgb=0, gce=1,
                 confi=0):
        super().__init__()
        self.eikonal_weight, self.rgb_weight = eikonal_weight, rgb_weight
        self.mvs_weight, self.sparse_weight = mvs_weight, sparse_weight
        self.gce, self.anneal_rgb, self.confi = gce, anneal_rgb, confi
        self.rgb_loss = _class_by_path(rgb_loss)(reduction='mean')
        self.iter_step = 0

    def set_stg(self, stg):
        # loss.py:31-36: later stages are not implemented by the reference either
        self.iter_step = 0
        if stg >= 1:
            self.anneal_rgb = 0
            self.sparse_weight = 0
            raise NotImplementedError

    # ---- terms ---------------------------------------------------------------------------------------
    @staticmethod
    def _conf_ray(model_outputs):
        """sum_s p_i p_j per ray: from the fused lookup (`CostMapper.mvs_loss`) when it ran, else from pi / pj"""
        if 'conf_ray' in model_outputs:
            return model_outputs['conf_ray']
        return (model_outputs['pi'] * model_outputs['pj']).sum(-1)

    def get_rgb_loss(self, rgb_values, rgb_gt, model_outputs=None, t=0):
        rgb_gt = rgb_gt.reshape(-1, 3)
        if t > 0:   # only rays the MVS volumes are uncertain about (loss.py:40-45)
            uncertain = self._conf_ray(model_outputs) < t
            return ((rgb_values - rgb_gt).abs().mean(-1) * uncertain).mean()
        return self.rgb_loss(rgb_values, rgb_gt)

    def get_eikonal_loss(self, grad_theta):
        return ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()

    def get_mvs_loss(self, model_outputs):
        if 'mvs_loss_fused' in model_outputs:   # lookup + this term in one kernel (svs_mvs_loss), same gce / confi
            return model_outputs['mvs_loss_fused']
        pw = model_outputs['pi'] * model_outputs['pj']
        w = model_outputs['weights']
        if self.gce == 1:
            per_sample = -pw * w
        elif self.gce == 0:
            per_sample = -pw * torch.log(w + 1e-8)
        else:
            per_sample = -pw * w.detach() ** self.gce * torch.log(w + 1e-8)
        confident = (pw.sum(1) > self.confi).to(per_sample.dtype)
        return (confident * per_sample.sum(1)).mean()

    def get_sparse_loss(self, model_outputs):
        conf_ray = self._conf_ray(model_outputs)
        key = 'depth_values_all' if 'depth_values_all' in model_outputs else 'depth_values'
        return ((1. / (model_outputs[key].squeeze() + 1e-3)) * (conf_ray < self.confi)).mean()

    # ---- total ---------------------------------------------------------------------------------------
    def forward_device(self, model_outputs, ground_truth, iter_step):
        """forward() with the iteration counter as a DEVICE scalar and no host-side branch on it: the annealing switch
        (`iter_step < anneal_rgb`, loss.py:99-108) becomes a blend of both photometric terms, so the whole loss can be
        captured in a CUDA graph and replayed across the switch (svolsdf_b200.train.GraphedTrainStep increments the
        counter inside the graph).  Same values as forward() for every iteration."""
        dev = model_outputs['rgb_values'].device
        zero = torch.zeros((), device=dev)
        has_mvs = 'pi' in model_outputs or 'mvs_loss_fused' in model_outputs
        out = {
            'rgb_loss': self.get_rgb_loss(model_outputs['rgb_values'], ground_truth['rgb'].to(dev)),
            'eikonal_loss': self.get_eikonal_loss(model_outputs['grad_theta']) if 'grad_theta' in model_outputs else zero,
            'mvs_loss': self.get_mvs_loss(model_outputs) if has_mvs and self.mvs_weight > 0 else zero,
            'sparse_loss': zero,
        }
        sparse_term = zero
        if has_mvs and self.sparse_weight > 0 and self.anneal_rgb > 0:
            it = iter_step.to(torch.float32)
            ann = (it < self.anneal_rgb).to(torch.float32)
            smooth = self.get_rgb_loss(model_outputs['rgb_values'], ground_truth['rgb_smooth'].to(dev), model_outputs, t=1e-8)
            out['rgb_loss'] = ann * smooth + (1.0 - ann) * out['rgb_loss']
            out['sparse_loss'] = ann * self.get_sparse_loss(model_outputs)
            anneal_sparse = ann * (1.0 - (it / self.anneal_rgb).clamp(0.0, 1.0))
            sparse_term = self.sparse_weight * anneal_sparse * out['sparse_loss']
        out['loss'] = self.rgb_weight * out['rgb_loss'] + self.eikonal_weight * out['eikonal_loss'] + \
            self.mvs_weight * out['mvs_loss'] + sparse_term
        return out

    def forward(self, model_outputs, ground_truth):
        dev = model_outputs['rgb_values'].device
        zero = torch.zeros((), device=dev)
        rgb_gt = ground_truth['rgb'].to(dev)
        has_mvs = 'pi' in model_outputs or 'mvs_loss_fused' in model_outputs
        annealing = self.sparse_weight > 0 and self.anneal_rgb > 0 and self.iter_step < self.anneal_rgb
        out = {
            'rgb_loss': self.get_rgb_loss(model_outputs['rgb_values'], rgb_gt),
            'eikonal_loss': self.get_eikonal_loss(model_outputs['grad_theta']) if 'grad_theta' in model_outputs else zero,
            'mvs_loss': self.get_mvs_loss(model_outputs) if has_mvs and self.mvs_weight > 0 else zero,
            'sparse_loss': self.get_sparse_loss(model_outputs) if has_mvs and annealing else zero,
        }
        anneal_sparse = 0
        if annealing:   # linear 1 -> 0 over anneal_rgb steps; the photometric term then only sees uncertain rays
            anneal_sparse = 1.0 - min(max(self.iter_step / self.anneal_rgb, 0.0), 1.0)
            out['rgb_loss'] = self.get_rgb_loss(model_outputs['rgb_values'], ground_truth['rgb_smooth'].to(dev),
                                                model_outputs, t=1e-8)
        out['loss'] = self.rgb_weight * out['rgb_loss'] + self.eikonal_weight * out['eikonal_loss'] + \
            self.mvs_weight * out['mvs_loss'] + self.sparse_weight * anneal_sparse * out['sparse_loss']
        self.iter_step += 1
        return out
