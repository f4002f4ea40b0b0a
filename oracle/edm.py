"""Oracle: the EDM-style ADM network ``DhariwalUNet`` (``--model_type adm`` WITHOUT ``--use_origin_adm``: the
ffhq_adm / bed_adm / imnet_adm presets), functional fp32 restatement (TEST INFRASTRUCTURE).

Follows reference ``models/EDM.py``:

* ``DhariwalUNet.__init__`` ``:716-810`` (module tree, names ``enc.{res}x{res}_{conv,down,block{i}}``,
  ``dec.{res}x{res}_{in0,in1,up,block{i}}``)                                         -> :func:`edm_plan`
* ``DhariwalUNet.forward`` ``:812-845`` / ``forward_with_cfg`` ``:847-861``          -> :func:`edm_forward` / ``_with_cfg``
* ``UNetBlock.forward`` ``:254-292``: ``conv0(silu(norm0(x)))``; ``silu(shift + norm1(x) * (scale + 1))`` with
  ``(scale, shift) = affine(emb).chunk(2)`` (scale FIRST); ``conv1``; ``+ skip(orig)``; ``* skip_scale`` (= 1 here);
  optional self-attention on the block's OUTPUT: channel index of ``qkv`` = ``head * 3 dh + c * 3 + {q,k,v}``
  (``reshape(B*heads, dh, 3, T).unbind(2)``, ``:277-281``), weights ``softmax(q . k / sqrt(dh))`` (``:160-170``)
* ``Conv2d.forward`` ``:101-134`` with ``resample_filter=[1, 1]``: ``down`` = 2x2 mean then the 3x3 conv,
  ``up`` = nearest x2 then the 3x3 conv; a ``kernel=0`` skip is the bare resampling (``:241-252``)
* ``GroupNorm`` ``:139-153``: ``min(32, C // 4)`` groups, eps 1e-5
* ``PositionalEmbedding`` ``:490-506``: ``[cos(t f), sin(t f)]``, ``f_i = (1/10000)^(i / (C/2))``, raw t
* label path ``:822-829``: ``emb += map_label(one_hot(y))`` (a column of the bias-free ``map_label.weight``);
  ``drop_half_label`` zeroes the one-hot of the second half of the batch (the CFG null class)

Pinned against the reference's own module by ``oracle/make_goldens.py`` (edm_* fixtures).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class EDMConfig:
    img_resolution: int = 32         # latent side (config.image_size // config.f)
    in_channels: int = 4
    out_channels: int = 4
    label_dim: int = 0               # 0 = unconditional
    model_channels: int = 256
    channel_mult: tuple = (1, 2, 3, 4)
    num_blocks: int = 2
    attn_resolutions: tuple = (16, 8, 4)   # feature-map RESOLUTIONS with self-attention (EDM.py:788), not ds rates
    channels_per_head: int = 64

    @property
    def emb_dim(self):
        return self.model_channels * 4


def edm_plan(cfg: EDMConfig):
    """Ordered module list.  Each entry: dict(name, kind, cin, cout, res, attn, up, down) with kind in
    {'conv', 'block'}; ``res`` is the resolution the block's OUTPUT lives at.  Returns (enc, dec, final_ch)."""
    mc = cfg.model_channels
    enc = []
    cout = cfg.in_channels
    for level, mult in enumerate(cfg.channel_mult):
        res = cfg.img_resolution >> level
        if level == 0:
            cin, cout = cout, mc * mult
            enc.append(dict(name=f"enc.{res}x{res}_conv", kind="conv", cin=cin, cout=cout, res=res))
        else:
            enc.append(dict(name=f"enc.{res}x{res}_down", kind="block", cin=cout, cout=cout, res=res, attn=False,
                            up=False, down=True))
        for idx in range(cfg.num_blocks):
            cin, cout = cout, mc * mult
            enc.append(dict(name=f"enc.{res}x{res}_block{idx}", kind="block", cin=cin, cout=cout, res=res,
                            attn=res in cfg.attn_resolutions, up=False, down=False))
    skips = [e["cout"] for e in enc]
    dec = []
    for level, mult in reversed(list(enumerate(cfg.channel_mult))):
        res = cfg.img_resolution >> level
        if level == len(cfg.channel_mult) - 1:
            dec.append(dict(name=f"dec.{res}x{res}_in0", kind="block", cin=cout, cout=cout, res=res, attn=True,
                            up=False, down=False))
            dec.append(dict(name=f"dec.{res}x{res}_in1", kind="block", cin=cout, cout=cout, res=res, attn=False,
                            up=False, down=False))
        else:
            dec.append(dict(name=f"dec.{res}x{res}_up", kind="block", cin=cout, cout=cout, res=res, attn=False,
                            up=True, down=False))
        for idx in range(cfg.num_blocks + 1):
            cin = cout + skips.pop()
            cout = mc * mult
            dec.append(dict(name=f"dec.{res}x{res}_block{idx}", kind="block", cin=cin, cout=cout, res=res,
                            attn=res in cfg.attn_resolutions, up=False, down=False))
    return enc, dec, cout


def param_shapes(cfg: EDMConfig) -> "dict[str, tuple]":
    """Reference ``state_dict()`` keys and shapes in registration order.  Includes the constant ``resample_filter``
    buffers ([1,1,2,2], all 0.25) that up/down ``Conv2d`` modules register (EDM.py:96-98)."""
    E, mc = cfg.emb_dim, cfg.model_channels
    s = {"map_layer0.weight": (E, mc), "map_layer0.bias": (E,), "map_layer1.weight": (E, E), "map_layer1.bias": (E,)}
    if cfg.label_dim:
        s["map_label.weight"] = (E, cfg.label_dim)
    enc, dec, ch = edm_plan(cfg)
    for m in enc + dec:
        p = m["name"] + "."
        if m["kind"] == "conv":
            s[p + "weight"] = (m["cout"], m["cin"], 3, 3)
            s[p + "bias"] = (m["cout"],)
            continue
        cin, cout = m["cin"], m["cout"]
        resample = m["up"] or m["down"]
        s[p + "norm0.weight"] = (cin,); s[p + "norm0.bias"] = (cin,)
        s[p + "conv0.weight"] = (cout, cin, 3, 3); s[p + "conv0.bias"] = (cout,)
        if resample:
            s[p + "conv0.resample_filter"] = (1, 1, 2, 2)
        s[p + "affine.weight"] = (2 * cout, E); s[p + "affine.bias"] = (2 * cout,)
        s[p + "norm1.weight"] = (cout,); s[p + "norm1.bias"] = (cout,)
        s[p + "conv1.weight"] = (cout, cout, 3, 3); s[p + "conv1.bias"] = (cout,)
        if cin != cout:
            s[p + "skip.weight"] = (cout, cin, 1, 1); s[p + "skip.bias"] = (cout,)
        if resample:
            s[p + "skip.resample_filter"] = (1, 1, 2, 2)
        if m["attn"]:
            s[p + "norm2.weight"] = (cout,); s[p + "norm2.bias"] = (cout,)
            s[p + "qkv.weight"] = (3 * cout, cout, 1, 1); s[p + "qkv.bias"] = (3 * cout,)
            s[p + "proj.weight"] = (cout, cout, 1, 1); s[p + "proj.bias"] = (cout,)
    s["out_norm.weight"] = (ch,); s["out_norm.bias"] = (ch,)
    s["out_conv.weight"] = (cfg.out_channels, ch, 3, 3); s["out_conv.bias"] = (cfg.out_channels,)
    return s


def synthetic_state_dict(cfg: EDMConfig, seed: int = 1) -> "dict[str, torch.Tensor]":
    """Seeded non-degenerate weights (the reference zero-inits conv1, proj and out_conv, so a fresh model returns 0).
    conv / linear: U(-a, a), a = 1/sqrt(fan_in); GroupNorm: weight 1 + 0.1 N, bias 0.1 N; biases 0.02 N;
    ``map_label.weight`` 0.02 N; ``resample_filter`` = 0.25 (its only legal value).  One CPU generator, keys in
    :func:`param_shapes` order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        leaf = k.rsplit(".", 2)[-2]
        if k.endswith("resample_filter"):
            sd[k] = torch.full(shp, 0.25)
        elif leaf.startswith("norm") or leaf == "out_norm":
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)) if k.endswith("weight") else 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias") or k == "map_label.weight":
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return sd


def positional_embedding(t: torch.Tensor, dim: int, max_positions: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.arange(0, half, dtype=torch.float32) / half
    freqs = (1 / max_positions) ** freqs
    args = torch.outer(t.float(), freqs)
    return torch.cat([args.cos(), args.sin()], dim=1)


def _gn(x, w, b):
    c = x.shape[1]
    return F.group_norm(x, min(32, c // 4), w, b, eps=1e-5)


def _down(x):
    return F.avg_pool2d(x, 2)


def _up(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _block(sd, m, x, emb, cfg):
    p = m["name"] + "."
    orig = x
    h = F.silu(_gn(x, sd[p + "norm0.weight"], sd[p + "norm0.bias"]))
    if m["up"]:
        h = _up(h)
    if m["down"]:
        h = _down(h)
    h = F.conv2d(h, sd[p + "conv0.weight"], sd[p + "conv0.bias"], padding=1)
    params = F.linear(emb, sd[p + "affine.weight"], sd[p + "affine.bias"])[:, :, None, None]
    scale, shift = params.chunk(2, dim=1)
    h = F.silu(shift + _gn(h, sd[p + "norm1.weight"], sd[p + "norm1.bias"]) * (scale + 1))
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    if m["up"]:
        orig = _up(orig)
    if m["down"]:
        orig = _down(orig)
    if p + "skip.weight" in sd:
        orig = F.conv2d(orig, sd[p + "skip.weight"], sd[p + "skip.bias"])
    x = h + orig
    if m["attn"]:
        b, c, hh, ww = x.shape
        heads = c // cfg.channels_per_head
        qkv = F.conv2d(_gn(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), sd[p + "qkv.weight"], sd[p + "qkv.bias"])
        q, k, v = qkv.reshape(b * heads, c // heads, 3, -1).unbind(2)
        w = torch.einsum("ncq,nck->nqk", q, k / math.sqrt(k.shape[1])).softmax(dim=2)
        a = torch.einsum("nqk,nck->ncq", w, v)
        x = F.conv2d(a.reshape(b, c, hh, ww), sd[p + "proj.weight"], sd[p + "proj.bias"]) + x
    return x


@torch.no_grad()
def edm_forward(sd, cfg: EDMConfig, t, x, y=None, drop_half_label: bool = False):
    """v = DhariwalUNet(noise_labels=t, x, y)  (EDM.py:812-845).  t: 0-d or [B]."""
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
    emb = positional_embedding(t, cfg.model_channels)
    emb = F.silu(F.linear(emb, sd["map_layer0.weight"], sd["map_layer0.bias"]))
    emb = F.linear(emb, sd["map_layer1.weight"], sd["map_layer1.bias"])
    if cfg.label_dim and y is not None:
        onehot = F.one_hot(torch.as_tensor(y), cfg.label_dim).float()
        if drop_half_label:
            onehot[len(onehot) // 2:] *= 0.0
        emb = emb + onehot @ sd["map_label.weight"].t()
    emb = F.silu(emb)
    enc, dec, _ = edm_plan(cfg)
    skips = []
    h = x.float()
    for m in enc:
        if m["kind"] == "conv":
            h = F.conv2d(h, sd[m["name"] + ".weight"], sd[m["name"] + ".bias"], padding=1)
        else:
            h = _block(sd, m, h, emb, cfg)
        skips.append(h)
    for m in dec:
        if h.shape[1] != m["cin"]:
            h = torch.cat([h, skips.pop()], dim=1)
        h = _block(sd, m, h, emb, cfg)
    h = F.silu(_gn(h, sd["out_norm.weight"], sd["out_norm.bias"]))
    return F.conv2d(h, sd["out_conv.weight"], sd["out_conv.bias"], padding=1)


@torch.no_grad()
def edm_forward_with_cfg(sd, cfg: EDMConfig, t, x, y, cfg_scale: float):
    """EDM.py:847-861: first half duplicated, labels of the second half dropped, ``u + s (c - u)`` in both halves."""
    half = x[: len(x) // 2]
    out = edm_forward(sd, cfg, t, torch.cat([half, half], 0), y, drop_half_label=True)
    cond, uncond = torch.split(out, len(out) // 2, dim=0)
    he = uncond + cfg_scale * (cond - uncond)
    return torch.cat([he, he], 0)


def edm_flops_per_sample(cfg: EDMConfig) -> int:
    """Algorithmic FLOPs (2 MAC) of the convolutions, linears and attention matmuls per sample per NFE."""
    enc, dec, chf = edm_plan(cfg)
    E = cfg.emb_dim
    fl = 2 * (cfg.model_channels * E + E * E)
    for m in enc + dec:
        px = m["res"] * m["res"]
        if m["kind"] == "conv":
            fl += 2 * px * 9 * m["cin"] * m["cout"]
            continue
        cin, cout = m["cin"], m["cout"]
        fl += 2 * px * 9 * cin * cout + 2 * px * 9 * cout * cout + 2 * E * 2 * cout
        if cin != cout:
            fl += 2 * px * cin * cout
        if m["attn"]:
            fl += 2 * px * cout * 3 * cout + 2 * px * cout * cout + 4 * px * px * cout
    fl += 2 * cfg.img_resolution ** 2 * 9 * chf * cfg.out_channels
    return fl
