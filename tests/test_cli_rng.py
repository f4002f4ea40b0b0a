"""CLI surface and seed generators (CPU): flags/defaults of test_flow_latent.py:302-408, and the generators
against streams recorded from the reference's sampler/random_util.py (tests/golden/rng_*.npz)."""
import numpy as np
import torch

from lfm_b200 import cli
from lfm_b200.random_util import get_generator
from tests._util import load_golden


def test_reference_flag_defaults():
    a = cli.build_parser().parse_args([])
    assert (a.generator, a.seed, a.n_sample, a.model_type) == ("determ", 42, 50000, "adm")
    assert (a.method, a.atol, a.rtol, a.step_size, a.num_steps) == ("dopri5", 1e-5, 1e-5, 0.01, 40)
    assert (a.scale_factor, a.f, a.batch_size, a.cfg_scale) == (0.18215, 8, 200, 1.0)
    assert not a.use_karras_samplers and not a.compute_fid and not a.compute_nfe and not a.measure_time
    b = cli.build_parser().parse_args("--model_type DiT-L/2 --image_size 256 --num_in_channels 4 --num_classes 1 "
                                      "--label_dropout 0. --use_karras_samplers --method heun --num_steps 25".split())
    assert b.model_type == "DiT-L/2" and b.use_karras_samplers and b.method == "heun" and b.num_steps == 25


def test_determ_generator_matches_reference_stream():
    g = load_golden("rng_determ")
    gen = get_generator("determ", 16, 42)
    x = gen.randn(4, 4, 32, 32)
    y = gen.randint(0, 10, (4,))
    assert np.array_equal(x.numpy(), g["x"]) and np.array_equal(y.numpy(), g["y"])


def test_determ_indiv_generator_matches_reference_stream():
    g = load_golden("rng_indiv")
    gen = get_generator("determ-indiv", 16, 42)
    assert np.array_equal(gen.randn(4, 4, 32, 32).numpy(), g["x"])


def test_determ_is_batch_size_independent():
    a = get_generator("determ", 32, 7).randn(8, 4, 4, 4)
    b = get_generator("determ", 32, 7).randn(3, 4, 4, 4)
    assert torch.equal(a[:3], b)


def test_cli_builds_every_native_network_family():
    """The reference command lines of the three network families (bash_scripts/run_test.sh with test_args/*.txt) parse
    and build through create_network with seeded weights - on the CPU, without running the network."""
    import torch
    from lfm_b200 import DhariwalUNet, DiT, UNetModel
    from lfm_b200.cli import build_parser, load_model
    common = ["--image_size", "256", "--f", "8", "--num_in_channels", "4", "--num_out_channels", "4", "--nf", "128",
              "--synthetic_init", "3", "--no_decode"]
    cases = [
        (["--model_type", "DiT-S/2", "--num_classes", "1", "--label_dropout", "0."], DiT),
        (["--model_type", "DiT-S/8", "--num_classes", "10", "--label_dropout", "0.1"], DiT),   # 4 x 4 token grid
        (["--model_type", "adm", "--use_origin_adm", "--ch_mult", "1", "2", "--attn_resolutions", "2", "--num_res_blocks", "1",
          "--num_heads", "2"], UNetModel),
        (["--model_type", "adm", "--ch_mult", "1", "2", "--attn_resolutions", "16", "--num_res_blocks", "1", "--label_dim", "7"],
         DhariwalUNet),
    ]
    for extra, cls in cases:
        args = build_parser().parse_args(common + extra)
        net = load_model(args, torch.device("cpu"))
        assert isinstance(net, cls)
        assert all(torch.isfinite(p).all() for p in net.state_dict().values())
        assert float(sum(p.abs().sum() for p in net.parameters())) > 0
