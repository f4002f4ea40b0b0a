"""Oracle: the OpenAI-ADM ``UNetModel`` velocity network, functional fp32 restatement (TEST INFRASTRUCTURE).

Follows reference ``models/guided_diffusion/unet.py``:

* ``UNetModel.__init__`` ``:407-595`` (block wiring)           -> :func:`unet_plan`
* ``UNetModel.forward`` ``:613-655``                              -> :func:`unet_forward`
* ``ResBlock._forward`` ``:218-238`` (use_scale_shift_norm: ``GN(h) * (1 + scale) + shift``, scale FIRST)
* ``AttentionBlock._forward`` ``:281-287`` + ``QKVAttentionLegacy.forward`` ``:319-334`` (channel index =
  head*3ch + {q,k,v}*ch + c; scale ch^-1/4 on q and k)
* ``Downsample`` ``:103-128`` (conv3x3 stride 2) / ``Upsample`` ``:73-100`` (nearest x2 then conv3x3)
* ``nn.py:17-19`` GroupNorm32(32, C) (eps 1e-5, affine), ``nn.py:103-121`` timestep_embedding (cos first, raw t)

Only the configuration the LFM presets use (``get_flow_model``, models/__init__.py:46-68 with
test_args/celeb*_adm.txt): ``use_scale_shift_norm=True``, ``resblock_updown=False``, ``conv_resample=True``, legacy
attention order.  Pinned against the reference's own module by ``oracle/make_goldens.py`` (unet_* fixtures).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetConfig:
    image_size: int = 64            # latent side (config.image_size // 8)
    in_channels: int = 4
    model_channels: int = 256
    out_channels: int = 4
    num_res_blocks: int = 2
    attention_resolutions: tuple = (16, 8)   # DOWNSAMPLE RATES ds at which attention is applied (unet.py:482)
    channel_mult: tuple = (1, 2, 2, 2, 4)
    num_heads: int = 4
    num_head_channels: int = -1
    num_classes: int | None = None

    @property
    def emb_dim(self):
        return self.model_channels * 4


def unet_plan(cfg: UNetConfig):
    """The module tree as a list of blocks; each block is a list of layer tuples
    ('conv_in', cin, cout) | ('res', cin, cout) | ('attn', ch, heads) | ('down', ch) | ('up', ch).
    Returns (input_blocks, middle, output_blocks, final_ch).  Mirrors unet.py:460-590."""
    mc = cfg.model_channels

    def heads(ch):
        return cfg.num_heads if cfg.num_head_channels == -1 else ch // cfg.num_head_channels

    ch = int(cfg.channel_mult[0] * mc)
    inputs = [[("conv_in", cfg.in_channels, ch)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, int(mult * mc))]
            ch = int(mult * mc)
            if ds in cfg.attention_resolutions:
                layers.append(("attn", ch, heads(ch)))
            inputs.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inputs.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("attn", ch, heads(ch)), ("res", ch, ch)]
    outputs = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, int(mc * mult))]
            ch = int(mc * mult)
            if ds in cfg.attention_resolutions:
                layers.append(("attn", ch, heads(ch)))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs, ch


def param_shapes(cfg: UNetConfig) -> "dict[str, tuple]":
    """Reference state_dict keys and shapes, in registration order (celeb512 preset: 396 tensors)."""
    E = cfg.emb_dim
    s = {"time_embed.0.weight": (E, cfg.model_channels), "time_embed.0.bias": (E,),
         "time_embed.2.weight": (E, E), "time_embed.2.bias": (E,)}
    if cfg.num_classes is not None:
        s["label_emb.weight"] = (cfg.num_classes, E)

    def add_layers(prefix, layers):
        for j, L in enumerate(layers):
            p = f"{prefix}.{j}."
            if L[0] == "conv_in":
                s[p + "weight"] = (L[2], L[1], 3, 3); s[p + "bias"] = (L[2],)
            elif L[0] == "res":
                cin, cout = L[1], L[2]
                s[p + "in_layers.0.weight"] = (cin,); s[p + "in_layers.0.bias"] = (cin,)
                s[p + "in_layers.2.weight"] = (cout, cin, 3, 3); s[p + "in_layers.2.bias"] = (cout,)
                s[p + "emb_layers.1.weight"] = (2 * cout, E); s[p + "emb_layers.1.bias"] = (2 * cout,)
                s[p + "out_layers.0.weight"] = (cout,); s[p + "out_layers.0.bias"] = (cout,)
                s[p + "out_layers.3.weight"] = (cout, cout, 3, 3); s[p + "out_layers.3.bias"] = (cout,)
                if cin != cout:
                    s[p + "skip_connection.weight"] = (cout, cin, 1, 1); s[p + "skip_connection.bias"] = (cout,)
            elif L[0] == "attn":
                c = L[1]
                s[p + "norm.weight"] = (c,); s[p + "norm.bias"] = (c,)
                s[p + "qkv.weight"] = (3 * c, c, 1); s[p + "qkv.bias"] = (3 * c,)
                s[p + "proj_out.weight"] = (c, c, 1); s[p + "proj_out.bias"] = (c,)
            elif L[0] == "down":
                s[p + "op.weight"] = (L[1], L[1], 3, 3); s[p + "op.bias"] = (L[1],)
            elif L[0] == "up":
                s[p + "conv.weight"] = (L[1], L[1], 3, 3); s[p + "conv.bias"] = (L[1],)

    inputs, middle, outputs, ch = unet_plan(cfg)
    for i, layers in enumerate(inputs):
        add_layers(f"input_blocks.{i}", layers)
    add_layers("middle_block", middle)
    for i, layers in enumerate(outputs):
        add_layers(f"output_blocks.{i}", layers)
    s["out.0.weight"] = (ch,); s["out.0.bias"] = (ch,)
    s["out.2.weight"] = (cfg.out_channels, ch, 3, 3); s["out.2.bias"] = (cfg.out_channels,)
    return s


def synthetic_state_dict(cfg: UNetConfig, seed: int = 1) -> "dict[str, torch.Tensor]":
    """Seeded non-degenerate weights (the reference zero-inits out_layers.3, proj_out and out.2, unet.py:198,276,594,
    so a fresh model returns 0).  conv / linear: U(-a, a), a = 1/sqrt(fan_in); GroupNorm: weight 1 + 0.1 N, bias
    0.1 N; biases 0.02 N; embeddings 0.02 N.  One CPU generator, keys in param_shapes order."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
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
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return sd


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, w, b):
    return F.group_norm(x.float(), 32, w, b, eps=1e-5)


# Optional emulation of the CUDA path's operand rounding (unet_forward(..., emulate_bf16=True)): the operands of every
# tensor-core contraction (3x3 / 1x1 convolutions, qkv / proj) are rounded to bf16, accumulation stays fp32.  Used ONLY
# to give the adaptive solver's oracle run the same noise floor in v as the native path (tests/test_gpu_parity.py,
# cfg4): the embedded error estimate of dopri5 at rtol = 1e-5 is dominated by that noise, so the step count of an
# fp32 run is not comparable.  The default (False) is the plain fp32 restatement that the fixtures pin.
_EMULATE_BF16 = False


def _r(t):
    return t.bfloat16().float() if _EMULATE_BF16 else t


def _conv2d(x, w, b, **kw):
    return F.conv2d(_r(x), _r(w), b, **kw)


def _conv1d(x, w, b):
    return F.conv1d(_r(x), _r(w), b)


def _res(sd, p, x, emb):
    h = _conv2d(F.silu(_gn(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"])),
                sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])[..., None, None]
    scale, shift = torch.chunk(e, 2, dim=1)            # unet.py:232 - scale first
    h = _gn(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"]) * (1 + scale) + shift
    h = _conv2d(F.silu(h), sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if p + "skip_connection.weight" in sd:
        x = _conv2d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


def _attn(sd, p, x, heads):
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = _r(_conv1d(_gn(xf, sd[p + "norm.weight"], sd[p + "norm.bias"]), sd[p + "qkv.weight"], sd[p + "qkv.bias"]))
    ch = c // heads
    q, k, v = qkv.reshape(b * heads, ch * 3, hh * ww).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, hh * ww)
    h = _conv1d(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def _run_layers(sd, prefix, layers, h, emb):
    for j, L in enumerate(layers):
        p = f"{prefix}.{j}."
        if L[0] == "conv_in":
            h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
        elif L[0] == "res":
            h = _res(sd, p, h, emb)
        elif L[0] == "attn":
            h = _attn(sd, p, h, L[2])
        elif L[0] == "down":
            h = _conv2d(h, sd[p + "op.weight"], sd[p + "op.bias"], stride=2, padding=1)
        elif L[0] == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = _conv2d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)
    return h


@torch.no_grad()
def unet_forward(sd, cfg: UNetConfig, t, x, y=None, emulate_bf16: bool = False):
    """v = UNetModel(t, x, y)  (unet.py:613-655).  t: 0-d (expanded to [B], :629-630) or [B]."""
    global _EMULATE_BF16
    if emulate_bf16:
        _EMULATE_BF16 = True
        try:
            return unet_forward(sd, cfg, t, x, y)
        finally:
            _EMULATE_BF16 = False
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
    if t.numel() != x.shape[0]:
        t = t * torch.ones(x.shape[0])
    assert (y is not None) == (cfg.num_classes is not None)
    e = timestep_embedding(t, cfg.model_channels)
    emb = F.linear(F.silu(F.linear(e, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if cfg.num_classes is not None:
        emb = emb + sd["label_emb.weight"][y]
    inputs, middle, outputs, _ = unet_plan(cfg)
    hs = []
    h = x.float()
    for i, layers in enumerate(inputs):
        h = _run_layers(sd, f"input_blocks.{i}", layers, h, emb)
        hs.append(h)
    h = _run_layers(sd, "middle_block", middle, h, emb)
    for i, layers in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}", layers, h, emb)
    h = F.silu(_gn(h, sd["out.0.weight"], sd["out.0.bias"]))
    return _conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def unet_flops_per_sample(cfg: UNetConfig) -> int:
    """Algorithmic FLOPs (2 MAC) of the convolutions, linears and attention matmuls per sample per NFE."""
    inputs, middle, outputs, chf = unet_plan(cfg)
    E = cfg.emb_dim
    fl = 2 * (cfg.model_channels * E + E * E)
    side = cfg.image_size

    def layer(L, side):
        px = side * side
        if L[0] == "conv_in":
            return 2 * px * 9 * L[1] * L[2], side
        if L[0] == "res":
            cin, cout = L[1], L[2]
            f = 2 * px * 9 * cin * cout + 2 * px * 9 * cout * cout + 2 * E * 2 * cout
            if cin != cout:
                f += 2 * px * cin * cout
            return f, side
        if L[0] == "attn":
            c = L[1]
            return 2 * px * c * 3 * c + 2 * px * c * c + 4 * px * px * c, side
        if L[0] == "down":
            return 2 * (px // 4) * 9 * L[1] * L[1], side // 2
        if L[0] == "up":
            return 2 * (px * 4) * 9 * L[1] * L[1], side * 2
        raise KeyError(L)

    for layers in inputs:
        for L in layers:
            f, side = layer(L, side)
            fl += f
    for L in middle:
        f, side = layer(L, side)
        fl += f
    for layers in outputs:
        for L in layers:
            f, side = layer(L, side)
            fl += f
    fl += 2 * side * side * 9 * chf * cfg.out_channels
    return fl
