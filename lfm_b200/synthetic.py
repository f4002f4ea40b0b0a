"""Seeded non-degenerate weights for benchmarks and smoke tests.

The reference zero-initialises every adaLN modulation layer and the output layer (models/DiT.py:219-228), so a
freshly constructed network returns exactly 0 and neither timing nor parity on it means anything.  This
generator fills a DiT ``state_dict`` with the reference's own scales where it has them (xavier-uniform Linear
weights, N(0, 0.02) embeddings) and N(0, 0.02) for everything the reference zeroes.  One CPU generator, keys in
state_dict order => identical tensors on every machine (tests/test_host_logic.py checks it against the oracle's
independent copy of the same recipe).
"""
from __future__ import annotations

import math

import torch


def synthetic_state_dict(net, seed: int = 1) -> "dict[str, torch.Tensor]":
    """``net`` is an lfm_b200.DiT (possibly on the meta device); returns CPU fp32 tensors for load_state_dict."""
    from .network import _pos_embed_2d
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, p in net.state_dict().items():
        shp = tuple(p.shape)
        if k == "pos_embed":
            sd[k] = _pos_embed_2d(net.hidden_size, net.img_resolution // net.patch_size)
        elif k.endswith(".bias") or "adaLN" in k or "embedding_table" in k or k.startswith("t_embedder") \
                or k.startswith("final_layer"):
            sd[k] = torch.randn(shp, generator=g) * 0.02
        else:
            fan_out, fan_in = shp[0], int(math.prod(shp[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
    return sd
