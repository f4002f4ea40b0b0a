"""The oracle restatement vs. fixtures produced by the UNMODIFIED reference (oracle/make_goldens.py).

This is what pins the oracle: tests/golden/*.npz hold outputs of /root/reference's own DiT.forward,
forward_with_cfg and karras_sample on seeded inputs; weights are regenerated from (config, seed).
fp32 on both sides; tolerance covers summation-order differences only.
"""
import numpy as np
import pytest
import torch

from oracle import dit as odit
from oracle import solvers as osol
from tests._util import T, cfg_from_golden, load_golden, oracle_model, rel_l2

TOL = 2e-5


# patch 4 / 8, other latent sides, head_dim 72 and a 1024-token grid (make_goldens.py patch)
GEOMETRY = ["mini_p4", "mini_p8", "mini_r64p4", "mini_r16p2", "mini_xl2", "mini_xl4", "mini_r64p2"]


@pytest.mark.parametrize("name", ["mini_uncond", "mini_cond", "mini_d384"] + GEOMETRY)
def test_forward_matches_reference(name):
    g = load_golden(name)
    cfg = cfg_from_golden(g)
    sd = odit.synthetic_state_dict(cfg, int(g["weight_seed"]))
    x = T(g["x"])
    v = odit.dit_forward(sd, cfg, T(g["t_scalar"]), x, None)
    assert rel_l2(v, g["v_scalar_ynone"]) < TOL
    v = odit.dit_forward(sd, cfg, T(g["t_vec"]), x, T(g["y"]))
    assert rel_l2(v, g["v_vec_y"]) < TOL
    assert float(np.abs(g["v_vec_y"]).mean()) > 1e-3  # synthetic init is non-degenerate


@pytest.mark.parametrize("name", ["mini_cond", "mini_d384", "mini_p4", "mini_r16p2", "mini_xl4"])
def test_cfg_matches_reference(name):
    g = load_golden(name)
    cfg = cfg_from_golden(g)
    sd = odit.synthetic_state_dict(cfg, int(g["weight_seed"]))
    x = T(g["x"])
    x2 = torch.cat([x, x], 0)
    v = odit.dit_forward_with_cfg(sd, cfg, torch.full((4,), 0.4), x2, T(g["y_cfg"]), 1.5)
    assert rel_l2(v, g["v_cfg_1p5"]) < TOL
    assert torch.equal(v[:2], v[2:])


@pytest.mark.parametrize("name", ["mini_uncond", "mini_cond", "mini_d384"] + GEOMETRY)
def test_fixed_step_samplers_match_reference(name):
    g = load_golden(name)
    cfg = cfg_from_golden(g)
    sd = odit.synthetic_state_dict(cfg, int(g["weight_seed"]))
    x = T(g["x"])
    if cfg.num_classes > 1:
        f = oracle_model(sd, cfg, T(g["y_cfg"]), 1.5)
        xs = torch.cat([x, x], 0)
    else:
        f = oracle_model(sd, cfg)
        xs = x
    assert rel_l2(osol.karras_sample(f, xs, 6, "euler"), g["euler6"]) < TOL
    assert rel_l2(osol.karras_sample(f, xs, 5, "heun"), g["heun5"]) < TOL


def test_heun_corrector_quirk_matches_reference():
    """sample_heun's guard uses the default steps=40 (karras_sample.py:129,155): with 43 nodes the
    intervals 39..41 are Euler-only.  The 'corrected everywhere' variant must differ."""
    g = load_golden("mini_uncond")
    cfg = cfg_from_golden(g)
    sd = odit.synthetic_state_dict(cfg, int(g["weight_seed"]))
    f = oracle_model(sd, cfg)
    x = T(g["x"])
    assert rel_l2(osol.karras_sample(f, x, 43, "heun"), g["heun43"]) < 5e-5
    assert rel_l2(osol.karras_sample(f, x, 43, "heun", corrector_limit=1000), g["heun43"]) > 1e-5


def test_full_size_dit_l2():
    g = load_golden("dit_l2")
    cfg = odit.make_config("DiT-L/2", num_classes=1, label_dropout=0.0)
    sd = odit.synthetic_state_dict(cfg, int(g["weight_seed"]))
    assert len(sd) == 252
    x = T(g["x"])
    assert rel_l2(odit.dit_forward(sd, cfg, T(g["t"]), x), g["v"]) < TOL
    assert rel_l2(osol.karras_sample(oracle_model(sd, cfg), x, 3, "euler"), g["euler3"]) < TOL


def test_full_size_dit_b2_cfg():
    g = load_golden("dit_b2")
    cfg = odit.make_config("DiT-B/2", num_classes=1000, label_dropout=0.1)
    sd = odit.synthetic_state_dict(cfg, int(g["weight_seed"]))
    x = T(g["x"])
    v = odit.dit_forward_with_cfg(sd, cfg, T(g["t"]), torch.cat([x, x], 0), T(g["y_cfg"]), 1.5)
    assert rel_l2(v, g["v_cfg_1p5"]) < TOL


def test_pos_embed_spots():
    g = load_golden("pos_embed_spots")
    pe = odit.pos_embed_2d(1024, 16)
    assert abs(float(pe[0, 1, 0]) - np.sin(1.0)) < 1e-6          # column goes first
    assert float(pe[0, 16, 0]) == 0.0
    assert np.array_equal(pe[0, 17].numpy(), g["row17"])


def test_flop_formula():
    assert odit.dit_flops_per_sample(odit.make_config("DiT-L/2")) == 161_386_856_448
    assert odit.dit_flops_per_sample(odit.make_config("DiT-B/2")) == 46_003_912_704


def test_eager_baseline_restatement_matches_reference_fixture():
    """tests/tools/eager_dit.py (the plain-PyTorch SDPA baseline bench.py times on the GPU) computes the reference's
    DiT: checked here on the CPU against the reference-generated fixture, so the baseline is a fair one."""
    from tests.tools import eager_dit
    g = load_golden("mini_cond")
    cfg = cfg_from_golden(g)
    sd = odit.synthetic_state_dict(cfg, int(g["weight_seed"]))
    arch = dict(depth=cfg.depth, hidden=cfg.hidden_size, heads=cfg.num_heads)
    v = eager_dit.dit_forward(sd, T(g["t_vec"]), T(g["x"]), T(g["y"]), **arch)
    assert rel_l2(v, g["v_vec_y"]) < 1e-5
    v = eager_dit.dit_forward(sd, T(g["t_scalar"]), T(g["x"]), None, **arch)
    assert rel_l2(v, g["v_scalar_ynone"]) < 1e-5
