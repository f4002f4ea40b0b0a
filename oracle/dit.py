"""Oracle: DiT velocity network, functional fp32 restatement (TEST INFRASTRUCTURE, see oracle/__init__).

Follows the reference file by file:

* ``models/DiT.py:252-272``  ``DiT.forward(t, x, y)``               -> :func:`dit_forward`
* ``models/DiT.py:274-290``  ``DiT.forward_with_cfg``                -> :func:`dit_forward_with_cfg`
* ``models/DiT.py:43-69``    ``TimestepEmbedder`` (cos first, raw t) -> :func:`timestep_embedding`
* ``models/DiT.py:72-104``   ``LabelEmbedder`` (eval: plain lookup)   -> inside :func:`conditioning`
* ``models/DiT.py:112-131``  ``DiTBlock`` adaLN-Zero, chunk(6) order  -> :func:`dit_block`
* ``models/DiT.py:134-149``  ``FinalLayer``                           -> :func:`final_layer`
* ``models/DiT.py:230-243``  ``unpatchify`` (nhwpqc -> nchpwq)        -> :func:`unpatchify`
* ``models/DiT.py:299-346``  2-D sin-cos table ("w goes first")       -> :func:`pos_embed_2d`
* timm ``PatchEmbed`` (conv k=p, s=p, flatten(2).transpose(1,2)), ``Attention`` (qkv Linear,
  reshape(B,N,3,H,dh).permute(2,0,3,1,4), softmax(q k^T dh^-0.5) v, proj) and ``Mlp``
  (fc1, GELU(tanh), fc2) are third-party (timm, unpinned in requirements.txt:6); their published
  semantics are restated in :func:`patch_embed`, :func:`attention`, :func:`mlp`.

Weights are a plain ``dict[str, Tensor]`` with exactly the reference's ``state_dict`` keys, so the
same dict loads into the reference module (strict=True) and into ``lfm_b200``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class DiTConfig:
    """Constructor arguments of reference ``DiT.__init__`` (models/DiT.py:157-169)."""

    img_resolution: int = 32
    patch_size: int = 2
    in_channels: int = 4
    hidden_size: int = 1024
    depth: int = 24
    num_heads: int = 16
    mlp_ratio: float = 4.0
    label_dropout: float = 0.0
    num_classes: int = 1

    @property
    def grid(self) -> int:
        return self.img_resolution // self.patch_size

    @property
    def num_tokens(self) -> int:
        return self.grid * self.grid

    @property
    def table_rows(self) -> int:  # models/DiT.py:79-81
        return self.num_classes + (1 if self.label_dropout > 0 else 0)

    @property
    def patch_dim(self) -> int:
        return self.patch_size * self.patch_size * self.in_channels


# models/DiT.py:355-415 (the size table); only depth / width / heads / patch differ.
DIT_PRESETS = {
    "DiT-XL/2": dict(depth=28, hidden_size=1152, patch_size=2, num_heads=16),
    "DiT-XL/4": dict(depth=28, hidden_size=1152, patch_size=4, num_heads=16),
    "DiT-XL/8": dict(depth=28, hidden_size=1152, patch_size=8, num_heads=16),
    "DiT-L/2": dict(depth=24, hidden_size=1024, patch_size=2, num_heads=16),
    "DiT-L/4": dict(depth=24, hidden_size=1024, patch_size=4, num_heads=16),
    "DiT-L/8": dict(depth=24, hidden_size=1024, patch_size=8, num_heads=16),
    "DiT-B/2": dict(depth=12, hidden_size=768, patch_size=2, num_heads=12),
    "DiT-B/4": dict(depth=12, hidden_size=768, patch_size=4, num_heads=12),
    "DiT-B/8": dict(depth=12, hidden_size=768, patch_size=8, num_heads=12),
    "DiT-S/2": dict(depth=12, hidden_size=384, patch_size=2, num_heads=6),
    "DiT-S/4": dict(depth=12, hidden_size=384, patch_size=4, num_heads=6),
    "DiT-S/8": dict(depth=12, hidden_size=384, patch_size=8, num_heads=6),
}


def make_config(model_type: str, img_resolution=32, in_channels=4, label_dropout=0.0, num_classes=1) -> DiTConfig:
    """``create_network`` for DiT types (models/__init__.py:12-17)."""
    return DiTConfig(img_resolution=img_resolution, in_channels=in_channels, label_dropout=label_dropout,
                     num_classes=num_classes, **DIT_PRESETS[model_type])


# ----------------------------------------------------------------------------------------------
# fixed tables


def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    # models/DiT.py:327-346: omega_k = 10000^(-k/(dim/2)) in float64, [sin | cos]
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def pos_embed_2d(dim: int, grid: int) -> torch.Tensor:
    """[1, grid*grid, dim] fp32.  First dim/2 features encode the COLUMN (w), last dim/2 the row
    (models/DiT.py:305-323: ``np.meshgrid(grid_w, grid_h)`` - "here w goes first")."""
    gh = np.arange(grid, dtype=np.float32)
    gw = np.arange(grid, dtype=np.float32)
    mesh = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid, grid)
    emb = np.concatenate([_sincos_1d(dim // 2, mesh[0]), _sincos_1d(dim // 2, mesh[1])], axis=1)
    return torch.from_numpy(emb).float().unsqueeze(0)


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0) -> torch.Tensor:
    """models/DiT.py:43-62: [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(max_period) i / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ----------------------------------------------------------------------------------------------
# optional bf16 cast points (mirrors where the CUDA path rounds MMA operands / stored activations)


def _r(x: torch.Tensor, emulate: bool) -> torch.Tensor:
    return x.to(torch.bfloat16).float() if emulate else x


def _linear(x, w, b, emulate):
    return F.linear(_r(x, emulate), _r(w, emulate), b)


# ----------------------------------------------------------------------------------------------
# layers


def patch_embed(sd, cfg: DiTConfig, x: torch.Tensor, emulate=False) -> torch.Tensor:
    """timm PatchEmbed: Conv2d(C, D, k=p, s=p) -> flatten(2).transpose(1, 2); + pos_embed
    (models/DiT.py:179, 261).  Patch vector order is (c, p, q); tokens are row-major (h, w)."""
    p, g = cfg.patch_size, cfg.grid
    B, C = x.shape[0], x.shape[1]
    patches = x.reshape(B, C, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * p * p)
    w = sd["x_embedder.proj.weight"].reshape(cfg.hidden_size, -1)
    return F.linear(patches, w, sd["x_embedder.proj.bias"]) + sd["pos_embed"]


def conditioning(sd, cfg: DiTConfig, t: torch.Tensor, y, batch: int, emulate=False) -> torch.Tensor:
    """c = t_embedder(t) + y_embedder(y)  (models/DiT.py:259-264).  A 0-d t becomes [1] and
    broadcasts (models/DiT.py:65-66); ``y is None`` selects the table's last row."""
    if t.dim() == 0:
        t = t[None]
    tf = timestep_embedding(t, 256)
    h = F.linear(tf, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    temb = F.linear(F.silu(h), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    if y is None:
        y = torch.full((batch,), cfg.table_rows - 1, dtype=torch.long)
    yemb = sd["y_embedder.embedding_table.weight"][y]
    return temb + yemb


def attention(sd, pre: str, cfg: DiTConfig, x: torch.Tensor, emulate=False) -> torch.Tensor:
    """timm Attention(dim, num_heads, qkv_bias=True): out-feature index of qkv is
    ``which*D + head*dh + d``."""
    B, N, D = x.shape
    H = cfg.num_heads
    dh = D // H
    qkv = _r(_linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"], emulate), emulate)
    qkv = qkv.reshape(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q * dh ** -0.5) @ k.transpose(-2, -1)
    if emulate:  # the CUDA path feeds un-normalised bf16 probabilities to the PV MMA, fp32 row sum
        e = torch.exp(s - s.amax(dim=-1, keepdim=True))
        o = (_r(e, True) @ v) / e.sum(dim=-1, keepdim=True)
    else:
        o = s.softmax(dim=-1) @ v
    o = _r(o.transpose(1, 2).reshape(B, N, D), emulate)
    return _linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"], emulate)


def mlp(sd, pre: str, x: torch.Tensor, emulate=False) -> torch.Tensor:
    """timm Mlp: fc2(GELU_tanh(fc1(x))) (models/DiT.py:122-124)."""
    h = F.gelu(_linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"], emulate), approximate="tanh")
    return _linear(_r(h, emulate), sd[pre + "fc2.weight"], sd[pre + "fc2.bias"], emulate)


def _ln_mod(x, shift, scale):
    # LayerNorm(no affine, eps 1e-6) then modulate (models/DiT.py:20-21, 119, 121)
    D = x.shape[-1]
    return F.layer_norm(x, (D,), eps=1e-6) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def dit_block(sd, i: int, cfg: DiTConfig, x, c, emulate=False):
    pre = f"blocks.{i}."
    mod = F.linear(F.silu(c), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=1)  # models/DiT.py:128 - this exact order
    x = x + g1.unsqueeze(1) * attention(sd, pre + "attn.", cfg, _ln_mod(x, sh1, sc1), emulate)
    x = x + g2.unsqueeze(1) * mlp(sd, pre + "mlp.", _ln_mod(x, sh2, sc2), emulate)
    return x


def final_layer(sd, cfg: DiTConfig, x, c):
    mod = F.linear(F.silu(c), sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    return F.linear(_ln_mod(x, shift, scale), sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])


def unpatchify(cfg: DiTConfig, x: torch.Tensor) -> torch.Tensor:
    """models/DiT.py:230-243: out-feature order (p, q, c) - channel LAST, unlike patch_embed."""
    p, g, c = cfg.patch_size, cfg.grid, cfg.in_channels
    x = x.reshape(x.shape[0], g, g, p, p, c)
    return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, g * p, g * p)


@torch.no_grad()
def dit_forward(sd, cfg: DiTConfig, t, x, y=None, emulate_bf16=False, return_tokens=False):
    """v = DiT(t, x, y)   (models/DiT.py:252-272).  t: 0-d or [B]; x: [B,C,H,W] fp32; y: [B] int64."""
    t = torch.as_tensor(t, dtype=torch.float32)
    B = x.shape[0]
    h = patch_embed(sd, cfg, x.float(), emulate_bf16)
    c = conditioning(sd, cfg, t, y, B, emulate_bf16)
    if c.shape[0] == 1 and B != 1:
        c = c.expand(B, -1)
    for i in range(cfg.depth):
        h = dit_block(sd, i, cfg, h, c, emulate_bf16)
    if return_tokens:
        return h
    return unpatchify(cfg, final_layer(sd, cfg, h, c))


@torch.no_grad()
def dit_forward_with_cfg(sd, cfg: DiTConfig, t, x, y, cfg_scale: float, emulate_bf16=False):
    """models/DiT.py:274-290: first half of x duplicated, one 2B forward, ``u + s (c - u)`` over all
    in_channels, result duplicated in both halves."""
    half = x[: len(x) // 2]
    out = dit_forward(sd, cfg, t, torch.cat([half, half], dim=0), y, emulate_bf16)
    cond, uncond = torch.split(out, len(out) // 2, dim=0)
    g = uncond + cfg_scale * (cond - uncond)
    return torch.cat([g, g], dim=0)


# ----------------------------------------------------------------------------------------------
# parameters


def param_shapes(cfg: DiTConfig) -> "dict[str, tuple]":
    """Reference ``state_dict`` keys and shapes in registration order; DiT-L/2: 252."""
    D, Hd = cfg.hidden_size, int(cfg.hidden_size * cfg.mlp_ratio)
    p, C = cfg.patch_size, cfg.in_channels
    s = {"pos_embed": (1, cfg.num_tokens, D),
         "x_embedder.proj.weight": (D, C, p, p), "x_embedder.proj.bias": (D,),
         "t_embedder.mlp.0.weight": (D, 256), "t_embedder.mlp.0.bias": (D,),
         "t_embedder.mlp.2.weight": (D, D), "t_embedder.mlp.2.bias": (D,),
         "y_embedder.embedding_table.weight": (cfg.table_rows, D)}
    for i in range(cfg.depth):
        b = f"blocks.{i}."
        s[b + "attn.qkv.weight"] = (3 * D, D); s[b + "attn.qkv.bias"] = (3 * D,)
        s[b + "attn.proj.weight"] = (D, D); s[b + "attn.proj.bias"] = (D,)
        s[b + "mlp.fc1.weight"] = (Hd, D); s[b + "mlp.fc1.bias"] = (Hd,)
        s[b + "mlp.fc2.weight"] = (D, Hd); s[b + "mlp.fc2.bias"] = (D,)
        s[b + "adaLN_modulation.1.weight"] = (6 * D, D); s[b + "adaLN_modulation.1.bias"] = (6 * D,)
    s["final_layer.linear.weight"] = (p * p * C, D); s["final_layer.linear.bias"] = (p * p * C,)
    s["final_layer.adaLN_modulation.1.weight"] = (2 * D, D); s["final_layer.adaLN_modulation.1.bias"] = (2 * D,)
    return s


def synthetic_state_dict(cfg: DiTConfig, seed: int = 1) -> "dict[str, torch.Tensor]":
    """Seeded NON-DEGENERATE weights for a DiT of shape ``cfg``.

    The reference zero-initialises every adaLN layer and the final linear (models/DiT.py:219-228),
    so a freshly constructed model outputs exactly 0 and parity on it proves nothing (SURVEY.md
    "thing 4").  This generator keeps the reference's scales where it has them (xavier-uniform for
    Linear weights, N(0, 0.02) for embeddings, models/DiT.py:193-217) and fills everything the
    reference zeroes (biases, adaLN, final layer) with N(0, 0.02).  One CPU ``torch.Generator``,
    keys visited in ``param_shapes`` order => the same dict on every machine.  ``pos_embed`` is the
    fixed table, not random.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k == "pos_embed":
            sd[k] = pos_embed_2d(cfg.hidden_size, cfg.grid)
        elif k.endswith(".bias") or "adaLN" in k or "embedding_table" in k or k.startswith("t_embedder") \
                or k.startswith("final_layer"):
            sd[k] = torch.randn(shp, generator=g) * 0.02
        else:  # xavier-uniform on the [out, in-flat] view
            fan_out, fan_in = shp[0], int(np.prod(shp[1:]))
            a = math.sqrt(6.0 / (fan_in + fan_out))
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
    return sd


def dit_flops_per_sample(cfg: DiTConfig) -> int:
    """Algorithmic FLOPs per sample per network evaluation (FLOP = 2 MAC)."""
    L, T, D, P = cfg.depth, cfg.num_tokens, cfg.hidden_size, cfg.patch_dim
    return L * (24 * T * D * D + 4 * T * T * D + 12 * D * D) + 2 * T * P * D + 2 * (256 * D + D * D) + 4 * D * D \
        + 2 * T * D * P
