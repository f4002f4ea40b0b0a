"""Throughput of the other BASELINE.json configs on ONE GPU (the headline line is bench.py; these are records for
DESIGN.md / profiles, not bench lines).  python scripts/bench_configs.py [cfg3 cfg4 cfg5 edm]"""
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lfm_b200  # noqa: E402
from lfm_b200.synthetic import synthetic_edm_state_dict, synthetic_state_dict, synthetic_unet_state_dict  # noqa: E402

dev = torch.device("cuda:0")


def build_dit(model_type, num_classes, label_dropout, max_batch):
    with torch.device("meta"):
        net = lfm_b200.DiT_models[model_type](img_resolution=32, in_channels=4, label_dropout=label_dropout,
                                              num_classes=num_classes, max_batch=max_batch)
    sd = synthetic_state_dict(net, 1)
    net = net.to_empty(device="cpu")
    net.load_state_dict(sd, strict=True)
    return net.to(dev)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cfg3():
    # DiT-B/2 imnet, CFG 1.5, Heun-25 (48 NFE x 2 rows per image): per-GPU share of the 8-GPU run = 32 images
    net = build_dit("DiT-B/2", 1000, 0.1, 64)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(32, 4, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 1000, (32,), generator=g).to(dev)
    mk = dict(y=torch.cat([y, torch.full((32,), 1000, device=dev)]), cfg_scale=1.5)
    ms = timeit(lambda: lfm_b200.karras_sample(net, torch.cat([x, x]), 25, clip_denoised=False, model_kwargs=mk, sigma_min=1e-5,
                                               sigma_max=1.0, sampler="heun"))
    nfe = net.last_stats["nfe"]
    fl = nfe * 64 * 46_003_912_704
    return {"config": "cfg3 DiT-B/2 CFG1.5 Heun-25, 32 img/GPU (64-row forward)", "ms": ms, "nfe": nfe, "images_per_s": 32 / ms * 1e3,
            "tflops": fl / ms / 1e9}


def cfg4():
    # ADM UNetModel celeb512 (64x64x4), dopri5 atol=rtol=1e-5, batch 32
    with torch.device("meta"):
        net = lfm_b200.UNetModel(image_size=64, in_channels=4, model_channels=256, out_channels=4, num_res_blocks=2,
                                 attention_resolutions=(16, 8), channel_mult=(1, 2, 2, 2, 4), num_heads=4,
                                 use_scale_shift_norm=True, max_batch=32)
    sd = synthetic_unet_state_dict(net, 1)
    net = net.to_empty(device="cpu")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    x = torch.randn(32, 4, 64, 64, generator=torch.Generator().manual_seed(0)).to(dev)
    args = types.SimpleNamespace(method="dopri5", atol=1e-5, rtol=1e-5, cfg_scale=1.0, compute_nfe=True)
    ms = timeit(lambda: lfm_b200.sample_from_model(net, x, {}, args), reps=1)
    s = net.last_stats
    return {"config": "cfg4 ADM UNet celeb512 dopri5 1e-5, batch 32", "ms": ms, **s, "images_per_s": 32 / ms * 1e3,
            "nfe_img_per_s": s["nfe"] * 32 / ms * 1e3, "tflops": s["nfe"] * 32 * 189.72e9 / ms / 1e9}


def cfg5():
    # DiT-L/2, Euler NFE sweep, batch 128
    net = build_dit("DiT-L/2", 1, 0.0, 128)
    x = torch.randn(128, 4, 32, 32, generator=torch.Generator().manual_seed(0)).to(dev)
    out = []
    for n in (10, 20, 50, 100):
        args = types.SimpleNamespace(method="euler", step_size=1.0 / n, perturb=False, cfg_scale=1.0, compute_nfe=False)
        ms = timeit(lambda: lfm_b200.sample_from_model(net, x, {}, args), reps=2)
        out.append({"config": f"cfg5 DiT-L/2 Euler-{n}, batch 128", "ms": ms, "nfe": net.last_stats["nfe"], "images_per_s": 128 / ms * 1e3,
                    "tflops": net.last_stats["nfe"] * 128 * 161_386_856_448 / ms / 1e9})
    return out


def edm():
    # imnet_adm preset (test_args/imnet_adm.txt): DhariwalUNet, 1000 classes, CFG 1.25, dopri5 atol=rtol=1e-5; 32 images
    # => 64-row forward per NFE
    with torch.device("meta"):
        net = lfm_b200.DhariwalUNet(img_resolution=32, in_channels=4, out_channels=4, label_dim=1000, model_channels=256,
                                    channel_mult=(1, 2, 3, 4), num_blocks=2, attn_resolutions=(16, 8, 4), max_batch=64)
    sd = synthetic_edm_state_dict(net, 1)
    net = net.to_empty(device="cpu")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(32, 4, 32, 32, generator=g).to(dev)
    y = torch.randint(0, 1000, (32,), generator=g).to(dev)
    mk = dict(y=torch.cat([y, torch.zeros_like(y)]), cfg_scale=1.25)
    args = types.SimpleNamespace(method="dopri5", atol=1e-5, rtol=1e-5, cfg_scale=1.25, compute_nfe=True)
    ms = timeit(lambda: lfm_b200.sample_from_model(net, torch.cat([x, x]), mk, args), reps=1)
    s = net.last_stats
    return {"config": "imnet_adm DhariwalUNet CFG1.25 dopri5 1e-5, 32 img (64-row forward)", "ms": ms, **s,
            "images_per_s": 32 / ms * 1e3, "nfe_img_per_s": s["nfe"] * 32 / ms * 1e3, "tflops": s["nfe"] * 64 * 73.05e9 / ms / 1e9}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5", "edm"]
    res = []
    for w in which:
        r = {"cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5, "edm": edm}[w]()
        res.extend(r if isinstance(r, list) else [r])
        torch.cuda.empty_cache()
    for r in res:
        print(json.dumps(r))
