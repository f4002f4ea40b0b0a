"""lfm_b200.fid against tests/golden/fid_ref.npz, which oracle/make_fid_goldens.py records from the
reference's own pytorch_fid package (inception.py network code, fid_score.py distance)."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from lfm_b200 import fid

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fid_ref.npz")


def _images():
    g = torch.Generator().manual_seed(7)          # oracle/make_fid_goldens.py:seeded_images
    return torch.rand(2, 3, 64, 64, generator=g)


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


def test_inception_features_match_reference_network_cpu():
    z = np.load(GOLD)
    net = fid.FIDInception(fid.synthetic_inception_state_dict(1))
    f = net(_images())
    assert f.shape == (2, 2048)
    assert _rel(f.numpy(), z["feat"]) < 1e-5      # BN folded in fp64, fp32 convolutions: 2e-7 measured


def test_inception_uint8_nhwc_input_is_the_same_path():
    net = fid.FIDInception(fid.synthetic_inception_state_dict(1))
    u8 = (torch.rand(2, 40, 48, 3, generator=torch.Generator().manual_seed(1)) * 255).to(torch.uint8)
    a = net(u8)
    b = net(u8.permute(0, 3, 1, 2).float() / 255.0)     # ToTensor (fid_score.py:147)
    assert torch.equal(a, b)


def test_inception_state_dict_contract():
    shapes = fid.inception_state_dict_shapes()
    assert len(shapes) == 94 * 5                         # 94 BasicConv2d units up to Mixed_7c
    assert shapes["Mixed_7c.branch3x3dbl_3b.conv.weight"] == (384, 384, 3, 1)
    assert shapes["Mixed_6b.branch7x7_2.conv.weight"] == (128, 128, 1, 7)
    sd = fid.synthetic_inception_state_dict(1)
    bad = dict(sd)
    del bad["Conv2d_1a_3x3.bn.running_var"]
    with pytest.raises(KeyError):
        fid.FIDInception(bad)
    bad = dict(sd)
    bad["Conv2d_1a_3x3.conv.weight"] = torch.zeros(32, 3, 5, 5)
    with pytest.raises(ValueError):
        fid.FIDInception(bad)
    with pytest.raises(RuntimeError):
        fid.FIDInception()(_images())
    with pytest.raises(FileNotFoundError):
        fid.FIDInception.from_file("/nonexistent/pt_inception.pth")
    with pytest.raises(ValueError):
        fid.FIDInception(sd)(torch.zeros(2, 64, 64, 4, dtype=torch.uint8))


def test_frechet_distance_matches_reference():
    z = np.load(GOLD)
    assert abs(fid.frechet_distance(z["mu1"], z["s1"], z["mu2"], z["s2"]) - float(z["fid"])) < 1e-9 * float(z["fid"])
    # singular first covariance: scipy's sqrtm goes complex and the reference keeps the real part (fid_score.py:214-219)
    assert abs(fid.frechet_distance(z["mu3"], z["s3"], z["mu2"], z["s2"]) - float(z["fid_rank"])) < 1e-6 * float(z["fid_rank"])
    assert abs(fid.frechet_distance(z["mu1"], z["s1"], z["mu1"], z["s1"])) < 1e-8
    with pytest.raises(ValueError):
        fid.frechet_distance(z["mu1"], z["s1"], z["mu2"][:10], z["s2"])


def test_statistics_match_numpy():
    rng = np.random.RandomState(0)
    x = rng.randn(300, 32) * 3 + 100.0                   # large mean: the fp64 sums must not cancel
    st = fid.FIDStatistics(32)
    for i in range(0, 300, 70):
        st.update(torch.from_numpy(x[i:i + 70]).float())
    mu, sigma = st.finalize()
    x32 = x.astype(np.float32).astype(np.float64)
    assert np.allclose(mu.numpy(), x32.mean(0), rtol=1e-12)
    assert np.allclose(sigma.numpy(), np.cov(x32, rowvar=False), rtol=1e-9, atol=1e-9)     # fid_score.py:224-226
    with pytest.raises(ValueError):
        fid.FIDStatistics(32).finalize()
    with pytest.raises(ValueError):
        st.update(torch.zeros(3, 31))


def test_load_statistics_formats(tmp_path):
    mu, sigma = np.arange(4.0), np.eye(4) * 2
    np.savez(tmp_path / "a.npz", mu=mu, sigma=sigma)
    np.save(tmp_path / "b.npy", {"mu": mu, "sigma": sigma}, allow_pickle=True)       # the reference's *_stat.npy layout
    for name in ("a.npz", "b.npy"):
        m, s = fid.load_statistics(str(tmp_path / name))
        assert np.array_equal(m, mu) and np.array_equal(s, sigma)


def _stats_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.from_numpy(np.random.RandomState(5).randn(40, 16)).float()
    st = fid.FIDStatistics(16)
    st.update(x[rank::world])                 # the ddp loop's interleaved ownership (test_flow_latent_ddp.py:138)
    st.all_reduce()
    mu, sigma = st.finalize()
    q.put((rank, st.n, mu.numpy(), sigma.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_statistics_all_reduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = np.random.RandomState(5).randn(40, 16).astype(np.float32).astype(np.float64)
    for _, n, mu, sigma in res:
        assert n == 40
        assert np.allclose(mu, x.mean(0), rtol=1e-12, atol=1e-14)
        assert np.allclose(sigma, np.cov(x, rowvar=False), rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_fid_on_device():
    z = np.load(GOLD)
    dev = torch.device("cuda:0")
    net = fid.FIDInception(fid.synthetic_inception_state_dict(1)).to(dev)
    f = net(_images().to(dev))
    assert _rel(f.cpu().numpy(), z["feat"]) < 1e-4            # cuDNN fp32 (TF32 off) vs the reference network on CPU
    d = fid.frechet_distance(torch.from_numpy(z["mu1"]).to(dev), z["s1"], z["mu2"], z["s2"])
    assert abs(d - float(z["fid"])) < 1e-8 * float(z["fid"])
    # accumulator: uint8 NHWC batches in, one number out; identical image sets => distance 0 against their own statistics
    u8 = (torch.rand(24, 64, 64, 3, generator=torch.Generator().manual_seed(2)) * 255).to(torch.uint8).to(dev)
    acc = fid.FIDAccumulator(net, dev, batch=10)
    acc.update(u8[:16])
    acc.update(u8[16:])
    mu, sigma = acc.stats.finalize()
    feats = net(u8).double().cpu().numpy()
    # (batches of 10 / 6 / 8 and one of 24 may take different cuDNN algorithms: fp32 round-off apart, not bitwise)
    assert np.allclose(mu.cpu().numpy(), feats.mean(0), rtol=1e-4, atol=1e-5)
    assert np.allclose(sigma.cpu().numpy(), np.cov(feats, rowvar=False), rtol=1e-3, atol=1e-5)
