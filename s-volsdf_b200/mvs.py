"""MVS cost lookup of the reference's training loop.

`CostMapper` is the drop-in for `VolOpt.cost_mapping` (volsdf/vsdf.py:382-452): same call form
`(z_vals, ts, xyz_raw) -> (results_cost_j, results_cost_mvs, valid_mask)`, same per-view state as
`VolOpt.get_mvs_input` keeps (`costs[i]`, `z_mvs[i]` of shape (1, Dz, H, W), the views' intrinsics and poses, the
image resolution).  All views are evaluated by ONE kernel launch (`svs_cost_mapping`, csrc/mvs.cu) instead of the
reference's ~35 elementwise launches and three `grid_sample` calls per view; nothing runs on the CPU and there is no
PyTorch fallback.
"""
import torch

from . import _lib as L


class CostMapper(object):
    def __init__(self, costs, z_mvs, intrinsics, poses, view_ids, img_res, inverse_depth=True):
        """costs / z_mvs: sequences of (1, Dz, H, W) tensors (self.costs / self.z_mvs, vsdf.py:364-379);
        intrinsics / poses: the views' (4, 4) K and camera-to-world matrices (train_dataset.intrinsics_all[id_k],
        pose_all[id_k]); view_ids: self.trains_i; img_res: (height, width) of train_dataset.img_res;
        inverse_depth: hparams.inverse_depth (stage 0)."""
        assert len(costs) == len(z_mvs) == len(intrinsics) == len(poses) == len(view_ids)
        if len(costs) > 8:
            raise L.SvsError('svs_cost_mapping supports up to 8 source views (got %d)' % len(costs))
        self.view_ids = [int(v) for v in view_ids]
        self.img_res = (int(img_res[0]), int(img_res[1]))
        self.inverse_depth = bool(inverse_depth)
        self._vol = []          # device tensors kept alive: (cost, z_near, z_far) per view
        self._desc = []
        for c, z, K, P in zip(costs, z_mvs, intrinsics, poses):
            c = c.detach().reshape(c.shape[-3:]).float().contiguous().cuda()
            z = z.detach().reshape(z.shape[-3:]).float().cuda()
            zn, zf = z[0].contiguous(), z[-1].contiguous()
            self._vol.append((c, zn, zf))
            K, P = K.detach().float().cpu(), P.detach().float().cpu()
            v = L.MvsView()
            v.cost, v.z_near, v.z_far = L.ptr(c), L.ptr(zn), L.ptr(zf)
            v.Dz, v.H, v.W = int(c.shape[0]), int(c.shape[1]), int(c.shape[2])
            v.fx, v.fy, v.cx, v.cy, v.sk = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), float(K[0, 1])
            for r in range(3):
                for k in range(4):
                    v.c2w[4 * r + k] = float(P[r, k])
            self._desc.append(v)

    def __call__(self, z_vals, ts, xyz_raw):
        return self.cost_mapping(z_vals, ts, xyz_raw)

    def _views(self, ts):
        # a CUDA int32 tensor keeps the choice of the own view on the device (CUDA-graph replays with changing batches,
        # svolsdf_b200.train.GraphedTrainStep); anything else is read on the host like the reference's `ts[0] == id_k`
        own_dev = ts if (torch.is_tensor(ts) and ts.is_cuda and ts.dtype == torch.int32) else None
        own = -1 if own_dev is not None else (int(ts[0]) if torch.is_tensor(ts) else int(ts))
        arr = (L.MvsView * len(self._desc))()
        for i, v in enumerate(self._desc):
            arr[i] = v
            arr[i].view_id = self.view_ids[i]
            arr[i].same_view = 1 if self.view_ids[i] == own else 0
        return arr, own_dev

    def mvs_loss(self, weights, ts, xyz_raw, gce=1, confi=0):
        """Cost lookup FUSED with the MVS term of the loss (`VolSDFLoss.get_mvs_loss`, volsdf/model/loss.py:53-67): one
        kernel (`svs_mvs_loss`) reads the volumes, forms p_i p_j in registers and returns
          mvs_loss  () = mean over rays of [sum_s p_i p_j > confi] * sum_s term(p_i p_j, weights)   (differentiable in `weights`)
          conf_ray (N,) = sum_s p_i p_j   (no gradient; the loss's sparsity / uncertain-ray terms branch on it)
        — p_i, p_j never reach HBM and the backward is a multiply of the stored d term / d weights."""
        return _MvsLossFn.apply(self, weights, ts, xyz_raw, float(gce), float(confi))

    @torch.no_grad()
    def cost_mapping(self, z_vals, ts, xyz_raw):
        """z_vals (N, D) is only the shape / device template the reference uses it as; ts: batch image index
        (`ts[0] == id_k` marks the batch's own view, vsdf.py:392); xyz_raw (N, D, 3) world points."""
        xyz = xyz_raw.detach().float().contiguous()
        N, D = int(xyz.shape[0]), int(xyz.shape[1])
        arr, own_dev = self._views(ts)
        dev = xyz.device
        cost_j = torch.empty(N, D, dtype=torch.float32, device=dev)
        cost_mvs = torch.empty(N, D, dtype=torch.float32, device=dev)
        valid = torch.empty(N, D, dtype=torch.uint8, device=dev)
        L.call('svs_cost_mapping', L.ptr(xyz), N, D, arr, len(self._desc), self.img_res[0], self.img_res[1],
               1 if self.inverse_depth else 0, L.ptr(own_dev), L.ptr(cost_j), L.ptr(cost_mvs), L.ptr(valid), L.stream())
        return cost_j, cost_mvs, valid.bool()


class _MvsLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mapper, weights, ts, xyz_raw, gce, confi):
        xyz = xyz_raw.detach().float().contiguous()
        w = weights.detach().float().contiguous()
        N, D = int(xyz.shape[0]), int(xyz.shape[1])
        if tuple(w.shape) != (N, D):
            raise L.SvsError('mvs_loss: weights %s do not match the samples (%d, %d)' % (tuple(w.shape), N, D))
        arr, own_dev = mapper._views(ts)
        dev = xyz.device
        ray_loss = torch.empty(N, dtype=torch.float32, device=dev)
        conf_ray = torch.empty(N, dtype=torch.float32, device=dev)
        d_w = torch.empty(N, D, dtype=torch.float32, device=dev)
        L.call('svs_mvs_loss', L.ptr(xyz), N, D, arr, len(mapper._desc), mapper.img_res[0], mapper.img_res[1],
               1 if mapper.inverse_depth else 0, L.ptr(own_dev), L.ptr(w), gce, confi, L.ptr(ray_loss), L.ptr(conf_ray),
               L.ptr(d_w), L.stream())
        ctx.save_for_backward(d_w)
        ctx.n_rays = max(N, 1)
        ctx.mark_non_differentiable(conf_ray)
        return ray_loss.mean(), conf_ray

    @staticmethod
    def backward(ctx, g_loss, g_conf):
        (d_w,) = ctx.saved_tensors
        return None, d_w * (g_loss / ctx.n_rays), None, None, None, None
