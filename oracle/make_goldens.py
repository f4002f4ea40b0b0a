"""Generate tests/golden/*.npz from the UNMODIFIED reference modules.  Build-container only.

Run:  python oracle/make_goldens.py            (needs /root/reference; never run on the GPU box)

It imports ``/root/reference/models/DiT.py`` and ``/root/reference/sampler/karras_sample.py``
through ``oracle/timm_shim`` (timm is absent from the image), loads the seeded synthetic weights of
``oracle.dit.synthetic_state_dict`` into the reference ``DiT`` with ``strict=True`` (which also pins
the state-dict key set and shapes), runs the reference's own ``forward`` / ``forward_with_cfg`` /
``karras_sample`` on seeded inputs and stores inputs + outputs.  Weights are NOT stored: every test
regenerates them from (config, seed).  The committed fixtures are what pins the oracle (and, on the
GPU box, the CUDA path) to the reference's code.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LFM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "timm_shim"))
sys.path.insert(0, REF)

from oracle import dit as odit  # noqa: E402


def _import_reference():
    # models/__init__.py imports EDM + guided_diffusion too; those import cleanly with the shim.
    import importlib

    models = importlib.import_module("models")
    ks = importlib.import_module("sampler.karras_sample")
    return models, ks


CASES = {
    # name: (model kwargs for the reference DiT ctor, weight seed)
    "mini_uncond": (dict(depth=2, hidden_size=256, patch_size=2, num_heads=4, img_resolution=32, in_channels=4,
                         label_dropout=0.0, num_classes=1), 11),
    "mini_cond": (dict(depth=2, hidden_size=256, patch_size=2, num_heads=4, img_resolution=32, in_channels=4,
                       label_dropout=0.1, num_classes=10), 12),
    "mini_d384": (dict(depth=3, hidden_size=384, patch_size=2, num_heads=6, img_resolution=32, in_channels=4,
                       label_dropout=0.1, num_classes=5), 13),
}
# Other DiT geometries of the reference's size table (models/DiT.py:355-415: the /4 and /8 entries) and other latent sides
# (img_resolution = image_size // f): token grids of 8 x 8, 4 x 4 and - with patch 4 on 64 x 64 latents - 16 x 16.
# Written by `make_goldens.py patch` with its own generator, so the fixtures above stay byte-identical.
GEOMETRY_CASES = {
    "mini_p4": (dict(depth=2, hidden_size=256, patch_size=4, num_heads=4, img_resolution=32, in_channels=4,
                     label_dropout=0.1, num_classes=10), 21),
    "mini_p8": (dict(depth=2, hidden_size=384, patch_size=8, num_heads=6, img_resolution=32, in_channels=4,
                     label_dropout=0.0, num_classes=1), 22),
    "mini_r64p4": (dict(depth=2, hidden_size=256, patch_size=4, num_heads=4, img_resolution=64, in_channels=4,
                        label_dropout=0.0, num_classes=1), 23),
    "mini_r16p2": (dict(depth=2, hidden_size=256, patch_size=2, num_heads=4, img_resolution=16, in_channels=4,
                        label_dropout=0.1, num_classes=5), 24),
    # head_dim 72 (the DiT-XL width, two blocks) at 256 and 64 tokens, and a 1024-token grid (64 x 64 latents, patch 2)
    "mini_xl2": (dict(depth=2, hidden_size=1152, patch_size=2, num_heads=16, img_resolution=32, in_channels=4,
                      label_dropout=0.0, num_classes=1), 25),
    "mini_xl4": (dict(depth=2, hidden_size=1152, patch_size=4, num_heads=16, img_resolution=32, in_channels=4,
                      label_dropout=0.1, num_classes=7), 26),
    "mini_r64p2": (dict(depth=2, hidden_size=256, patch_size=2, num_heads=4, img_resolution=64, in_channels=4,
                        label_dropout=0.0, num_classes=1), 27),
}
FULL = {
    "dit_l2": ("DiT-L/2", dict(img_resolution=32, in_channels=4, label_dropout=0.0, num_classes=1), 1),
    "dit_b2": ("DiT-B/2", dict(img_resolution=32, in_channels=4, label_dropout=0.1, num_classes=1000), 1),
}


def build(models, kwargs, seed, factory=None):
    from models.DiT import DiT, DiT_models

    net = DiT_models[factory](**kwargs) if factory else DiT(**kwargs)
    cfg = odit.DiTConfig(img_resolution=net.x_embedder.img_size[0], patch_size=net.patch_size,
                         in_channels=net.in_channels, hidden_size=net.pos_embed.shape[-1], depth=len(net.blocks),
                         num_heads=net.num_heads, label_dropout=kwargs["label_dropout"],
                         num_classes=kwargs["num_classes"])
    sd = odit.synthetic_state_dict(cfg, seed)
    # the fixed table must agree with the reference's own buffer before we overwrite it
    assert torch.allclose(sd["pos_embed"], net.pos_embed.data, atol=0, rtol=0), "pos_embed restatement differs"
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net, cfg


UNET_CASES = {
    # small ADM nets with the celeb512 topology rules (attention at ds 2 and 4 => 16x16 / 8x8 tokens here)
    # (channel counts are multiples of 128, the native path's GroupNorm vector width)
    "unet_mini": (dict(image_size=32, in_channels=4, model_channels=128, out_channels=4, num_res_blocks=1,
                       attention_resolutions=(2, 4), channel_mult=(1, 2, 2), num_heads=2, num_head_channels=-1,
                       num_classes=None), 21, 2),
    "unet_mini_cond": (dict(image_size=32, in_channels=4, model_channels=128, out_channels=4, num_res_blocks=2,
                            attention_resolutions=(4, 8), channel_mult=(1, 1, 2, 3), num_heads=4, num_head_channels=-1,
                            num_classes=7), 22, 3),
}


def make_unet_goldens(g):
    """ADM UNetModel fixtures from the reference's own module (models/guided_diffusion/unet.py)."""
    from models.guided_diffusion.unet import UNetModel
    from oracle import unet as ounet

    for name, (kw, seed, B) in UNET_CASES.items():
        cfg = ounet.UNetConfig(**kw)
        net = UNetModel(image_size=kw["image_size"], in_channels=kw["in_channels"], model_channels=kw["model_channels"],
                        out_channels=kw["out_channels"], num_res_blocks=kw["num_res_blocks"],
                        attention_resolutions=kw["attention_resolutions"], dropout=0.0, channel_mult=kw["channel_mult"],
                        conv_resample=True, dims=2, num_classes=kw["num_classes"], use_checkpoint=False, use_fp16=False,
                        num_heads=kw["num_heads"], num_head_channels=kw["num_head_channels"], num_heads_upsample=-1,
                        use_scale_shift_norm=True, resblock_updown=False, use_new_attention_order=False)
        sd = ounet.synthetic_state_dict(cfg, seed)
        net.load_state_dict(sd, strict=True)   # pins the key set / shapes of oracle.unet.param_shapes
        net.eval()
        x = torch.randn(B, 4, kw["image_size"], kw["image_size"], generator=g)
        tv = torch.tensor([0.85, 0.3, 0.55][:B])
        out = {"x": x, "t_vec": tv, "weight_seed": np.int64(seed)}
        for k, v in kw.items():
            if v is not None and not isinstance(v, tuple):
                out["cfg_" + k] = np.float64(v)
        out["cfg_attention_resolutions"] = np.array(kw["attention_resolutions"], dtype=np.int64)
        out["cfg_channel_mult"] = np.array(kw["channel_mult"], dtype=np.int64)
        if kw["num_classes"] is None:
            out["v"] = net(tv, x)
        else:
            y = torch.randint(0, kw["num_classes"], (B,), generator=g)
            out["y"] = y
            out["v"] = net(tv, x, y)
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
        print("wrote", name, "absmean(v)=%.4f" % float(out["v"].abs().mean()))

    # the celeb256 preset at full width (test_args/celeb256_adm.txt): B=1
    kw = dict(image_size=32, in_channels=4, model_channels=256, out_channels=4, num_res_blocks=2,
              attention_resolutions=(16, 8), channel_mult=(1, 2, 2, 2), num_heads=4, num_head_channels=-1, num_classes=None)
    cfg = ounet.UNetConfig(**kw)
    net = UNetModel(image_size=32, in_channels=4, model_channels=256, out_channels=4, num_res_blocks=2,
                    attention_resolutions=(16, 8), dropout=0.0, channel_mult=(1, 2, 2, 2), conv_resample=True, dims=2,
                    num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=4, num_head_channels=-1,
                    num_heads_upsample=-1, use_scale_shift_norm=True, resblock_updown=False, use_new_attention_order=False)
    sd = ounet.synthetic_state_dict(cfg, 1)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = torch.randn(1, 4, 32, 32, generator=g)
    tv = torch.tensor([0.6])
    np.savez_compressed(os.path.join(OUT, "unet_celeb256.npz"), x=x.numpy(), t_vec=tv.numpy(), v=net(tv, x).numpy(),
                        weight_seed=np.int64(1), n_tensors=np.int64(len(sd)))
    print("wrote unet_celeb256", len(sd), "tensors")


EDM_CASES = {
    # DhariwalUNet (models/EDM.py) with the preset topology rules; channel counts are multiples of 128
    "edm_mini": (dict(img_resolution=16, label_dim=0, model_channels=128, channel_mult=(1, 2), num_blocks=1,
                      attn_resolutions=(8,)), 31, 2),
    "edm_mini_cond": (dict(img_resolution=32, label_dim=10, model_channels=128, channel_mult=(1, 2, 3), num_blocks=2,
                           attn_resolutions=(16, 8)), 32, 3),
    # the ffhq_adm / bed_adm preset at full width (test_args/ffhq_adm.txt: nf 256, ch_mult 1 2 3 4, attn 16 8 4)
    "edm_ffhq": (dict(img_resolution=32, label_dim=0, model_channels=256, channel_mult=(1, 2, 3, 4), num_blocks=2,
                      attn_resolutions=(16, 8, 4)), 1, 1),
    # self-attention on the 32 x 32 grid (1024 tokens; the constructor's default attn_resolutions start at 32, EDM.py:723)
    "edm_attn32": (dict(img_resolution=32, label_dim=0, model_channels=128, channel_mult=(1, 2), num_blocks=1,
                        attn_resolutions=(32, 16)), 33, 2),
}


def make_edm_goldens(g):
    """DhariwalUNet fixtures from the reference's own module (models/EDM.py:716-861)."""
    from models.EDM import DhariwalUNet
    from oracle import edm as oedm

    for name, (kw, seed, B) in EDM_CASES.items():
        cfg = oedm.EDMConfig(**kw)
        net = DhariwalUNet(img_resolution=kw["img_resolution"], in_channels=4, out_channels=4, label_dim=kw["label_dim"],
                           augment_dim=0, model_channels=kw["model_channels"], channel_mult=list(kw["channel_mult"]),
                           channel_mult_emb=4, num_blocks=kw["num_blocks"], attn_resolutions=list(kw["attn_resolutions"]),
                           dropout=0.0, label_dropout=0.0)
        ref_keys = list(net.state_dict().keys())
        assert ref_keys == list(oedm.param_shapes(cfg).keys()), "state_dict key order differs from oracle.edm.param_shapes"
        for k, v in net.state_dict().items():
            if k.endswith("resample_filter"):
                assert torch.all(v == 0.25) and tuple(v.shape) == (1, 1, 2, 2)
        sd = oedm.synthetic_state_dict(cfg, seed)
        net.load_state_dict(sd, strict=True)
        net.eval()
        S = kw["img_resolution"]
        x = torch.randn(B, 4, S, S, generator=g)
        tv = torch.tensor([0.85, 0.3, 0.55][:B])
        t0 = torch.tensor(0.42)
        out = {"x": x, "t_vec": tv, "t_scalar": t0, "weight_seed": np.int64(seed), "n_tensors": np.int64(len(sd)),
               "cfg_img_resolution": np.int64(S), "cfg_label_dim": np.int64(kw["label_dim"]),
               "cfg_model_channels": np.int64(kw["model_channels"]), "cfg_num_blocks": np.int64(kw["num_blocks"]),
               "cfg_channel_mult": np.array(kw["channel_mult"], dtype=np.int64),
               "cfg_attn_resolutions": np.array(kw["attn_resolutions"], dtype=np.int64)}
        out["v_scalar"] = net(t0, x)
        if kw["label_dim"]:
            y = torch.randint(0, kw["label_dim"], (B,), generator=g)
            out["y"] = y
            out["v"] = net(tv, x, y)
            # forward_with_cfg on the doubled batch; the second-half labels are dropped inside (drop_half_label), so
            # they only need to be legal class ids for one_hot
            x2 = torch.cat([x, x], 0)
            y2 = torch.cat([y, torch.zeros(B, dtype=torch.long)], 0)
            out["y_cfg"] = y2
            out["t_cfg"] = torch.tensor([0.4] * (2 * B))
            out["v_cfg_1p25"] = net.forward_with_cfg(out["t_cfg"], x2, y2, cfg_scale=1.25)
            # the reference's own fixed-step samplers on this network, with its CFG denoiser dispatch
            # (karras_sample.py:42-49 -> DhariwalUNet.forward_with_cfg) and without labels' second half mattering
            ks = sys.modules["sampler.karras_sample"]
            common = dict(device="cpu", clip_denoised=False, sigma_min=1e-5, sigma_max=1.0, s_tmin=0.0, s_tmax=1.0, s_churn=0.0)
            out["cfg_euler4"] = ks.karras_sample(net, x2, steps=4, sampler="euler", model_kwargs=dict(y=y2, cfg_scale=1.25), **common)
            out["cfg_heun3"] = ks.karras_sample(net, x2, steps=3, sampler="heun", model_kwargs=dict(y=y2, cfg_scale=1.25), **common)
            out["y_euler3"] = ks.karras_sample(net, x, steps=3, sampler="euler", model_kwargs=dict(y=y), **common)
        else:
            out["v"] = net(tv, x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
        print("wrote", name, len(sd), "tensors, absmean(v)=%.4f" % float(out["v"].abs().mean()))
        del net


def make_geometry_goldens(models, ks, g):
    """One evaluation (0-d t without labels, [B] t with labels, forward_with_cfg) and the reference's own Euler / Heun
    samplers on the reference DiT for each entry of GEOMETRY_CASES."""
    for name, (kw, seed) in GEOMETRY_CASES.items():
        net, cfg = build(models, kw, seed)
        B, S = 2, kw["img_resolution"]
        x = torch.randn(B, 4, S, S, generator=g)
        out = {"x": x, "weight_seed": np.int64(seed)}
        for k, v in kw.items():
            out["cfg_" + k] = np.float64(v)
        out["t_scalar"] = torch.tensor(0.61)
        out["v_scalar_ynone"] = net(out["t_scalar"], x)
        tv = torch.tensor([0.85, 0.2])
        y = torch.randint(0, kw["num_classes"], (B,), generator=g)
        out["t_vec"], out["y"] = tv, y
        out["v_vec_y"] = net(tv, x, y)
        if kw["num_classes"] > 1:
            x2 = torch.cat([x, x], 0)
            y2 = torch.cat([y, torch.full((B,), kw["num_classes"])], 0)
            out["y_cfg"] = y2
            out["v_cfg_1p5"] = net.forward_with_cfg(torch.tensor([0.4] * (2 * B)), x2, y2, cfg_scale=1.5)
            mk, xs = dict(y=y2, cfg_scale=1.5), x2
        else:
            mk, xs = {}, x
        common = dict(model_kwargs=mk, device="cpu", clip_denoised=False, sigma_min=1e-5, sigma_max=1.0,
                      s_tmin=0.0, s_tmax=1.0, s_churn=0.0)
        out["euler6"] = ks.karras_sample(net, xs, steps=6, sampler="euler", **common)
        out["heun5"] = ks.karras_sample(net, xs, steps=5, sampler="heun", **common)
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
        print("wrote", name, {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    models, ks = _import_reference()
    if sys.argv[1:] == ["patch"]:   # only the other-geometry DiT fixtures (own generator)
        make_geometry_goldens(models, ks, torch.Generator().manual_seed(2468))
        return
    if sys.argv[1:] == ["edm"]:     # only the DhariwalUNet fixtures (own generator: the others stay byte-identical)
        make_edm_goldens(torch.Generator().manual_seed(8765))
        return
    g = torch.Generator().manual_seed(1234)

    for name, (kw, seed) in CASES.items():
        net, cfg = build(models, kw, seed)
        B = 2
        x = torch.randn(B, 4, 32, 32, generator=g)
        out = {"x": x, "weight_seed": np.int64(seed)}
        for k, v in kw.items():
            out["cfg_" + k] = np.float64(v)
        # (a) scalar (0-d) t, y = None
        t0 = torch.tensor(0.73)
        out["t_scalar"] = t0
        out["v_scalar_ynone"] = net(t0, x)
        # (b) vector t, explicit labels
        tv = torch.tensor([0.9, 0.15])
        y = torch.randint(0, kw["num_classes"], (B,), generator=g)
        out["t_vec"], out["y"] = tv, y
        out["v_vec_y"] = net(tv, x, y)
        if kw["num_classes"] > 1:
            # (c) forward_with_cfg on the doubled batch (test_flow_latent.py:171-181)
            x2 = torch.cat([x, x], 0)
            y2 = torch.cat([y, torch.full((B,), kw["num_classes"])], 0)
            out["y_cfg"] = y2
            out["v_cfg_1p5"] = net.forward_with_cfg(torch.tensor([0.4] * (2 * B)), x2, y2, cfg_scale=1.5)
            mk = dict(y=y2, cfg_scale=1.5)
            xs = x2
        else:
            mk = {}
            xs = x
        # (d) the reference's own fixed-step samplers
        common = dict(model_kwargs=mk, device="cpu", clip_denoised=False, sigma_min=1e-5, sigma_max=1.0,
                      s_tmin=0.0, s_tmax=1.0, s_churn=0.0)
        out["euler6"] = ks.karras_sample(net, xs, steps=6, sampler="euler", **common)
        out["heun5"] = ks.karras_sample(net, xs, steps=5, sampler="heun", **common)
        if name == "mini_uncond":
            out["heun43"] = ks.karras_sample(net, xs, steps=43, sampler="heun", **common)  # corrector stops at i=39
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
        print("wrote", name, {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})

    for name, (factory, kw, seed) in FULL.items():
        net, cfg = build(models, kw, seed, factory)
        B = 2
        x = torch.randn(B, 4, 32, 32, generator=g)
        out = {"x": x, "weight_seed": np.int64(seed)}
        if kw["num_classes"] == 1:
            out["t"] = torch.tensor(0.5)
            out["v"] = net(out["t"], x)
            out["euler3"] = ks.karras_sample(net, x, steps=3, sampler="euler", model_kwargs={}, device="cpu",
                                             clip_denoised=False, sigma_min=1e-5, sigma_max=1.0)
        else:
            y = torch.randint(0, 1000, (B,), generator=g)
            y2 = torch.cat([y, torch.full((B,), 1000)], 0)
            out["t"] = torch.tensor([0.8] * (2 * B))
            out["y_cfg"] = y2
            out["v_cfg_1p5"] = net.forward_with_cfg(out["t"], torch.cat([x, x], 0), y2, cfg_scale=1.5)
        np.savez_compressed(os.path.join(OUT, name + ".npz"),
                            **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
        print("wrote", name)
        del net

    make_unet_goldens(torch.Generator().manual_seed(4321))
    make_edm_goldens(torch.Generator().manual_seed(8765))

    # spot values of the fixed table quoted in SURVEY.md 8(c)
    pe = odit.pos_embed_2d(1024, 16)
    np.savez_compressed(os.path.join(OUT, "pos_embed_spots.npz"), pe_0_1_0=pe[0, 1, 0].numpy(),
                        pe_0_16_0=pe[0, 16, 0].numpy(), row17=pe[0, 17].numpy())


if __name__ == "__main__":
    main()
