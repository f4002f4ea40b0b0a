"""Record tests/golden/rng_*.npz from the UNMODIFIED reference generators.  Build-container only.

Run:  python oracle/make_rng_goldens.py [--check]     (needs /root/reference; never run on the GPU box)

Imports ``/root/reference/sampler/random_util.py`` and records the streams tests/test_cli_rng.py pins
``lfm_b200.random_util`` against:

* ``rng_determ.npz``: ``get_generator("determ", num_samples=16, seed=42)`` -> ``randn(4, 4, 32, 32)`` then
  ``randint(0, 10, (4,))`` (DeterministicGenerator, random_util.py:36-96; CPU RNG => identical on every machine);
* ``rng_indiv.npz``:  ``get_generator("determ-indiv", 16, 42)`` -> ``randn(4, 4, 32, 32)``
  (DeterministicIndividualGenerator, random_util.py:99-173).

Single process, CPU: with ``torch.distributed`` uninitialised the reference generators fall back to rank 0 / world 1
(random_util.py:43-49).  ``--check`` compares against the committed files instead of writing.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LFM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def record():
    sys.path.insert(0, REF)
    from sampler import random_util as ru
    out = {}
    gen = ru.get_generator("determ", 16, 42)
    x = gen.randn(4, 4, 32, 32)
    y = gen.randint(0, 10, (4,))
    out["rng_determ"] = dict(x=x.cpu().numpy(), y=y.cpu().numpy())
    gen = ru.get_generator("determ-indiv", 16, 42)
    out["rng_indiv"] = dict(x=gen.randn(4, 4, 32, 32).cpu().numpy())
    return out


def main():
    check = "--check" in sys.argv
    rec = record()
    ok = True
    for name, d in rec.items():
        path = os.path.join(OUT, name + ".npz")
        if check:
            z = np.load(path)
            same = sorted(z.files) == sorted(d) and all(np.array_equal(z[k], d[k]) for k in d)
            print(f"{name}: {'identical' if same else 'DIFFERENT'}")
            ok &= same
        else:
            np.savez(path, **d)
            print("wrote", path)
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
