"""oracle.edm (DhariwalUNet restatement) vs fixtures produced by the UNMODIFIED reference module
(models/EDM.py via ``python oracle/make_goldens.py edm``)."""
import numpy as np
import pytest
import torch

from oracle import edm as oedm
from oracle import solvers as osol
from tests._util import T, load_golden, rel_l2


def edm_cfg_from_golden(g):
    return oedm.EDMConfig(img_resolution=int(g["cfg_img_resolution"]), label_dim=int(g["cfg_label_dim"]),
                          model_channels=int(g["cfg_model_channels"]), num_blocks=int(g["cfg_num_blocks"]),
                          channel_mult=tuple(int(v) for v in g["cfg_channel_mult"]),
                          attn_resolutions=tuple(int(v) for v in g["cfg_attn_resolutions"]))


@pytest.mark.parametrize("name", ["edm_mini", "edm_mini_cond", "edm_attn32"])
def test_edm_forward_matches_reference(name):
    g = load_golden(name)
    cfg = edm_cfg_from_golden(g)
    sd = oedm.synthetic_state_dict(cfg, int(g["weight_seed"]))
    assert len(sd) == int(g["n_tensors"])
    y = T(g["y"]) if "y" in g else None
    assert rel_l2(oedm.edm_forward(sd, cfg, T(g["t_vec"]), T(g["x"]), y), g["v"]) < 2e-5
    # 0-d t, y = None (for a class-conditional net the label term is simply skipped, EDM.py:823)
    assert rel_l2(oedm.edm_forward(sd, cfg, T(g["t_scalar"]), T(g["x"])), g["v_scalar"]) < 2e-5
    assert float(np.abs(g["v"]).mean()) > 1e-2
    if "v_cfg_1p25" in g:
        x2 = torch.cat([T(g["x"]), T(g["x"])], 0)
        v = oedm.edm_forward_with_cfg(sd, cfg, T(g["t_cfg"]), x2, T(g["y_cfg"]), 1.25)
        assert rel_l2(v, g["v_cfg_1p25"]) < 2e-5
        n = len(v) // 2
        assert torch.equal(v[:n], v[n:])


def test_edm_ffhq_preset():
    g = load_golden("edm_ffhq")
    cfg = oedm.EDMConfig()   # ffhq_adm / bed_adm preset
    assert edm_cfg_from_golden(g) == cfg
    sd = oedm.synthetic_state_dict(cfg, 1)
    assert len(sd) == int(g["n_tensors"]) == 428
    assert rel_l2(oedm.edm_forward(sd, cfg, T(g["t_vec"]), T(g["x"])), g["v"]) < 2e-5


def test_edm_plan_preset():
    cfg = oedm.EDMConfig()
    enc, dec, ch = oedm.edm_plan(cfg)
    assert [m["name"] for m in enc][:4] == ["enc.32x32_conv", "enc.32x32_block0", "enc.32x32_block1", "enc.16x16_down"]
    assert len(enc) == 12 and len(dec) == 17 and ch == 256
    # attention: 2 blocks per encoder level at 16/8/4, dec in0 + 3 blocks per decoder level at 4/8/16
    assert sum(m.get("attn", False) for m in enc) == 6 and sum(m["attn"] for m in dec) == 10
    # widest GroupNorm input is a concatenation (1024 + 1024)
    assert max(m["cin"] for m in dec) == 2048
    fl = oedm.edm_flops_per_sample(cfg)
    assert 50e9 < fl < 200e9


def test_edm_samplers_match_reference():
    """oracle.solvers driven by oracle.edm vs the reference's own karras_sample on its own DhariwalUNet, including the
    CFG denoiser dispatch (karras_sample.py:42-49 -> forward_with_cfg, EDM.py:847-861)."""
    g = load_golden("edm_mini_cond")
    cfg = edm_cfg_from_golden(g)
    sd = oedm.synthetic_state_dict(cfg, int(g["weight_seed"]))
    x, y, y2 = T(g["x"]), T(g["y"]), T(g["y_cfg"])
    x2 = torch.cat([x, x], 0)
    fc = lambda tt, xx: oedm.edm_forward_with_cfg(sd, cfg, tt, xx, y2, 1.25)  # noqa: E731
    assert rel_l2(osol.karras_sample(fc, x2, 4, "euler"), g["cfg_euler4"]) < 2e-5
    assert rel_l2(osol.karras_sample(fc, x2, 3, "heun"), g["cfg_heun3"]) < 2e-5
    f = lambda tt, xx: oedm.edm_forward(sd, cfg, tt, xx, y)  # noqa: E731
    assert rel_l2(osol.karras_sample(f, x, 3, "euler"), g["y_euler3"]) < 2e-5
