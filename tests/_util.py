"""Shared helpers for the parity tests (oracle side)."""
import os

import numpy as np
import torch

from oracle import dit as odit

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def cfg_from_golden(g) -> odit.DiTConfig:
    return odit.DiTConfig(img_resolution=int(g["cfg_img_resolution"]), patch_size=int(g["cfg_patch_size"]),
                          in_channels=int(g["cfg_in_channels"]), hidden_size=int(g["cfg_hidden_size"]),
                          depth=int(g["cfg_depth"]), num_heads=int(g["cfg_num_heads"]),
                          label_dropout=float(g["cfg_label_dropout"]), num_classes=int(g["cfg_num_classes"]))


def T(a):
    return torch.from_numpy(np.asarray(a))


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def oracle_model(sd, cfg, y=None, cfg_scale=1.0, emulate_bf16=False):
    """f(t, x) with the reference's denoiser dispatch (karras_sample.py:42-49)."""
    if cfg_scale > 1.0:
        return lambda t, x: odit.dit_forward_with_cfg(sd, cfg, t, x, y, cfg_scale, emulate_bf16)
    return lambda t, x: odit.dit_forward(sd, cfg, t, x, y, emulate_bf16)
