"""Numerical check of the 'LayerNorm folded into the consumer GEMM's epilogue' identity proposed in DESIGN.md section 7 (CPU, fp32 / emulated bf16).

    y = (LN(x) * (1 + scale) + shift) W^T + b                      (models/DiT.py:20-21,129-130: what qkv / fc1 compute today)
      = r * ((x * (1 + scale)) W^T) - r * mu * u + v,   u = W (1 + scale),  v = W shift + b,  r = rstd(x), mu = mean(x)

Today the GEMM's A operand is bf16(LN(x) * (1 + scale) + shift); the folded form would feed bf16(x * (1 + scale)) - no mean subtraction before the
rounding - and apply r, mu in the epilogue.  This script measures both roundings against the fp32 result on the token stream of a synthetic
DiT-L/2 (oracle weights), block by block, to see whether the folded form loses accuracy when |mean| is not small against the row's spread.
usage: python tests/tools/ln_fold_numerics.py [blocks]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dit as odit  # noqa: E402  (analysis script: not part of the product path)


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def main():
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    torch.manual_seed(0)
    cfg = odit.make_config("DiT-L/2", num_classes=1, label_dropout=0.0)
    cfg = odit.DiTConfig(**{**cfg.__dict__, "depth": nblk})
    sd = odit.synthetic_state_dict(cfg, 1)
    x = torch.randn(2, 4, 32, 32)
    h = odit.patch_embed(sd, cfg, x)
    c = odit.conditioning(sd, cfg, torch.tensor(0.6), None, 2).expand(2, -1)
    print("block  |mean|/std (median, max)   rel err of the GEMM output: today (bf16 of LN-modulated x)   folded (bf16 of x*(1+scale))")
    for i in range(nblk):
        pre = f"blocks.{i}."
        mod = torch.nn.functional.linear(torch.nn.functional.silu(c), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
        sh, sc = mod.chunk(6, dim=1)[0:2]
        W, b = sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]
        mu = h.mean(-1, keepdim=True)
        var = ((h - mu) ** 2).mean(-1, keepdim=True)
        r = torch.rsqrt(var + 1e-6)
        a = (h - mu) * r * (1 + sc[:, None]) + sh[:, None]
        ref = a @ W.T + b
        today = bf(a) @ bf(W).T + b
        xt = bf(h * (1 + sc[:, None]))
        u = (bf(W) @ (1 + sc).T).T                       # [B, N]
        v = (bf(W) @ sh.T).T + b
        folded = r * (xt @ bf(W).T) - r * mu * u[:, None] + v[:, None]
        ratio = (mu.abs() / var.sqrt()).flatten()
        e1 = float((today - ref).norm() / ref.norm())
        e2 = float((folded - ref).norm() / ref.norm())
        print(f"{i:5d}  {float(ratio.median()):.3f}, {float(ratio.max()):.3f}                 {e1:.2e}                                   {e2:.2e}")
        h = odit.dit_block(sd, i, cfg, h, c)


if __name__ == "__main__":
    main()
