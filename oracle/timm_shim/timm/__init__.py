"""Analysis-only stand-in for the `timm` package (absent from this image; no network).

Only used by oracle/make_goldens.py, in the build container, to import /root/reference/models/DiT.py
unmodified.  It restates the published semantics of the three timm classes DiT.py imports
(models/DiT.py:17).  Never imported by the product.
"""
