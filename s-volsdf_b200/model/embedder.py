"""NeRF positional encoding with the reference's interface (volsdf/model/embedder.py:38-50).

`get_embedder(multires, input_dims)` returns (embed_fn, out_dim); embed_fn runs the CUDA PE kernel.  Inside
the networks the encoding is fused into the MLP kernels and this function is not called.
"""
import torch

from .. import _lib as L


class Embedder(object):
    def __init__(self, multires, input_dims=3):
        self.n_freqs = multires
        self.input_dims = input_dims
        self.out_dim = input_dims * (1 + 2 * multires)

    def embed(self, inputs):
        x = inputs.detach().reshape(-1, self.input_dims).contiguous().float()
        out = torch.empty(x.shape[0], self.out_dim, dtype=torch.float32, device=x.device)
        L.call('svs_embed', L.ptr(x), x.shape[0], self.input_dims, self.n_freqs, L.ptr(out), L.stream())
        return out.reshape(tuple(inputs.shape[:-1]) + (self.out_dim,))


def get_embedder(multires, input_dims=3):
    eo = Embedder(multires, input_dims)
    return eo.embed, eo.out_dim
