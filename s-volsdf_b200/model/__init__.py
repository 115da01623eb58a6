"""Mirror of the reference's `volsdf.model` package (network, network_bg, ray_sampler, density, embedder)."""
