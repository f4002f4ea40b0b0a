"""Record tests/golden/fid_ref.npz from the UNMODIFIED reference ``pytorch_fid`` package.  Build-container only.

Run:  python oracle/make_fid_goldens.py [--check]      (needs /root/reference; never run on the GPU box)

What it pins (tests/test_fid.py compares ``lfm_b200.fid`` with it):

* ``feat``: ``InceptionV3([3])(images)[0]`` of /root/reference/pytorch_fid/inception.py:24-163 (``fid_inception_v3``
  :182-201 with its four patched block classes) on two seeded 3 x 64 x 64 images in [0, 1], with the weight download
  (inception.py:199 - no network here) replaced by ``lfm_b200.fid.synthetic_inception_state_dict(1)``, a seeded
  dictionary both sides can rebuild.  The network code that runs is the reference's.
* ``fid``: ``calculate_frechet_distance`` (fid_score.py:178-228) on seeded 64-dimensional Gaussian statistics
  (``mu1, s1, mu2, s2`` are stored too), and ``fid_rank``: the same with a rank-deficient first covariance (more
  dimensions than samples, the case where ``sqrtm`` returns a complex matrix whose imaginary part the reference drops).
* ``stat_mu_head`` / ``stat_sigma_trace``: the head of ``mu`` and ``trace(sigma)`` of the reference's own
  ``pytorch_fid/celebahq_stat.npy`` as ``compute_statistics_of_path`` (fid_score.py:231-238) returns them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LFM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "fid_ref.npz")


def seeded_images():
    g = torch.Generator().manual_seed(7)
    return torch.rand(2, 3, 64, 64, generator=g)


def seeded_stats(dims=64, n1=500, n2=400, seed=3):
    rng = np.random.RandomState(seed)
    mix1, mix2 = rng.randn(dims, dims) / np.sqrt(dims), rng.randn(dims, dims) / np.sqrt(dims)
    a = rng.randn(n1, dims) @ mix1 + 0.1 * rng.randn(dims)
    b = rng.randn(n2, dims) @ mix2
    return a, b


def record():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    from lfm_b200.fid import synthetic_inception_state_dict
    from pytorch_fid import fid_score, inception

    sd = synthetic_inception_state_dict(1)
    full = inception._inception_v3(num_classes=1008, aux_logits=False, weights=None).state_dict()
    for k, v in full.items():          # fc.* and num_batches_tracked: present in the published file, unused by block 3
        sd.setdefault(k, torch.zeros_like(v))
    inception.load_state_dict_from_url = lambda *a, **k: sd
    torch.manual_seed(0)
    net = inception.InceptionV3([3]).eval()
    with torch.no_grad():
        feat = net(seeded_images())[0].squeeze(3).squeeze(2)

    a, b = seeded_stats()
    mu1, s1, mu2, s2 = a.mean(0), np.cov(a, rowvar=False), b.mean(0), np.cov(b, rowvar=False)
    fid = fid_score.calculate_frechet_distance(mu1, s1, mu2, s2)
    a_small = a[:40]                      # 40 samples, 64 dimensions: singular covariance
    mu3, s3 = a_small.mean(0), np.cov(a_small, rowvar=False)
    fid_rank = fid_score.calculate_frechet_distance(mu3, s3, mu2, s2)

    m, s = fid_score.compute_statistics_of_path(os.path.join(REF, "pytorch_fid", "celebahq_stat.npy"), None, 1, 2048, "cpu")
    return dict(feat=feat.numpy(), mu1=mu1, s1=s1, mu2=mu2, s2=s2, fid=np.float64(fid), mu3=mu3, s3=s3,
                fid_rank=np.float64(fid_rank), stat_mu_head=np.asarray(m[:8], dtype=np.float64),
                stat_sigma_trace=np.float64(np.trace(s)), stat_dims=np.int64(m.shape[0]))


def main():
    rec = record()
    if "--check" in sys.argv:
        z = np.load(OUT)
        ok = sorted(z.files) == sorted(rec)
        for k in rec:
            same = ok and np.allclose(z[k], rec[k], rtol=1e-6, atol=1e-7)
            print(f"{k}: {'same' if same else 'DIFFERENT'}")
            ok &= bool(same)
        return 0 if ok else 1
    np.savez(OUT, **rec)
    print("wrote", OUT, {k: np.asarray(v).shape for k, v in rec.items()})
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
