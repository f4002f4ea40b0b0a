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


def synthetic_unet_state_dict(net, seed: int = 1) -> "dict[str, torch.Tensor]":
    """Same idea for the ADM UNetModel (its zero_module convs make a fresh model return 0, unet.py:198,276,594):
    conv / linear U(-a, a) with a = 1/sqrt(fan_in); GroupNorm weight 1 + 0.1 N, bias 0.1 N; biases 0.02 N."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, p in net.state_dict().items():
        shp = tuple(p.shape)
        is_gn = (k.endswith("in_layers.0.weight") or k.endswith("out_layers.0.weight") or k.endswith("norm.weight")
                 or k == "out.0.weight")
        is_gn_b = (k.endswith("in_layers.0.bias") or k.endswith("out_layers.0.bias") or k.endswith("norm.bias")
                   or k == "out.0.bias")
        if is_gn:
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif is_gn_b:
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias") or k == "label_emb.weight":
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = int(math.prod(shp[1:]))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return sd


def synthetic_edm_state_dict(net, seed: int = 1) -> "dict[str, torch.Tensor]":
    """Same idea for the EDM-style DhariwalUNet (conv1 / proj / out_conv are zero-initialised, EDM.py:742): conv /
    linear U(-a, a) with a = 1/sqrt(fan_in); GroupNorm weight 1 + 0.1 N, bias 0.1 N; biases and map_label 0.02 N;
    resample_filter buffers keep their only legal value 0.25."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, p in net.state_dict().items():
        shp = tuple(p.shape)
        leaf = k.rsplit(".", 2)[-2]
        if k.endswith("resample_filter"):
            sd[k] = torch.full(shp, 0.25)
        elif leaf.startswith("norm") or leaf == "out_norm":
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)) if k.endswith("weight") else 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias") or k == "map_label.weight":
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = int(math.prod(shp[1:]))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return sd
