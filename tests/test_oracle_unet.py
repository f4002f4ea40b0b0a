"""oracle.unet (ADM UNetModel restatement) vs fixtures produced by the UNMODIFIED reference module
(models/guided_diffusion/unet.py via oracle/make_goldens.py)."""
import numpy as np
import pytest
import torch

from oracle import unet as ounet
from tests._util import T, load_golden, rel_l2


def unet_cfg_from_golden(g):
    return ounet.UNetConfig(image_size=int(g["cfg_image_size"]), in_channels=int(g["cfg_in_channels"]),
                            model_channels=int(g["cfg_model_channels"]), out_channels=int(g["cfg_out_channels"]),
                            num_res_blocks=int(g["cfg_num_res_blocks"]),
                            attention_resolutions=tuple(int(v) for v in g["cfg_attention_resolutions"]),
                            channel_mult=tuple(int(v) for v in g["cfg_channel_mult"]), num_heads=int(g["cfg_num_heads"]),
                            num_head_channels=int(g["cfg_num_head_channels"]),
                            num_classes=int(g["cfg_num_classes"]) if "cfg_num_classes" in g else None)


@pytest.mark.parametrize("name", ["unet_mini", "unet_mini_cond"])
def test_unet_forward_matches_reference(name):
    g = load_golden(name)
    cfg = unet_cfg_from_golden(g)
    sd = ounet.synthetic_state_dict(cfg, int(g["weight_seed"]))
    y = T(g["y"]) if "y" in g else None
    v = ounet.unet_forward(sd, cfg, T(g["t_vec"]), T(g["x"]), y)
    assert rel_l2(v, g["v"]) < 2e-5
    assert float(np.abs(g["v"]).mean()) > 1e-2


def test_unet_celeb256_preset():
    g = load_golden("unet_celeb256")
    cfg = ounet.UNetConfig(image_size=32, channel_mult=(1, 2, 2, 2))
    sd = ounet.synthetic_state_dict(cfg, 1)
    assert len(sd) == int(g["n_tensors"]) == 304
    v = ounet.unet_forward(sd, cfg, T(g["t_vec"]), T(g["x"]))
    assert rel_l2(v, g["v"]) < 2e-5
    # 0-d t is expanded to the batch (unet.py:629-630)
    v0 = ounet.unet_forward(sd, cfg, torch.tensor(0.6), T(g["x"]))
    assert rel_l2(v0, g["v"]) < 2e-5


def test_unet_plan_celeb512():
    cfg = ounet.UNetConfig()  # celeb512 preset (test_args/celeb512_adm.txt)
    assert len(ounet.param_shapes(cfg)) == 396
    inputs, middle, outputs, ch = ounet.unet_plan(cfg)
    assert len(inputs) == 15 and len(outputs) == 15 and ch == 256
    n_attn = sum(L[0] == "attn" for blk in inputs + [middle] + outputs for L in blk)
    assert n_attn == 11
    fl = ounet.unet_flops_per_sample(cfg)
    assert abs(fl / 189.72e9 - 1) < 0.01                             # 189.72 GFLOP / sample / NFE
    assert abs(ounet.unet_flops_per_sample(ounet.UNetConfig(image_size=32, channel_mult=(1, 2, 2, 2))) / 45.92e9 - 1) < 0.01
