"""CPU oracle for the LFM sampling hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (fp32, CPU) restatement of the reference's algorithm for the
sampling path: the DiT velocity network (reference ``models/DiT.py`` + the timm ``Attention`` /
``Mlp`` / ``PatchEmbed`` semantics it imports) and the ODE solvers around it (reference
``sampler/karras_sample.py`` and the torchdiffeq ``euler`` / ``dopri5`` paths called from
``test_flow_latent.py:42-76``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU baseline.  The product
(``lfm_b200``) never imports this package: it fails loudly when its CUDA library is missing.

Parity pin status
-----------------
* DiT forward, ``forward_with_cfg``, ``sample_euler`` / ``sample_heun`` / ``karras_sample``:
  PINNED against the reference's own modules, imported in the build container through
  ``oracle/timm_shim`` (``oracle/make_goldens.py`` regenerates ``tests/golden/*.npz`` from
  ``/root/reference``; ``tests/test_oracle_golden.py`` checks the oracle against them).
* timm (``Attention``, ``Mlp``, ``PatchEmbed``) and torchdiffeq (``odeint_adjoint``: fixed-grid
  euler, dopri5) are third-party dependencies that are absent from ``/root/reference`` and unpinned
  in its ``requirements.txt``; their published algorithms are restated here (timm 0.9.x,
  torchdiffeq 0.2.3).  For those two pieces the parity is UNPINNED by the reference: there are no
  reference tests or golden vectors for them; they are anchored on analytic known-answer tests
  (linear vector fields, NFE counts, scipy RK45).
"""
