"""Minimal stand-in for the nine `timm` symbols the reference hot-path files import.

TEST INFRASTRUCTURE ONLY.  `timm` is not installed in this image so the
reference modules under /root/reference cannot be imported without it.  This shim restates, from
timm 0.9.2's published behaviour, exactly the symbols used at
  GA/ga_convnext.py:15-19, GA/ga_cswin.py:14-17, MAP/models/map_convnext.py:4-6
so that `tests/golden/make_golden.py` can run the UNMODIFIED reference on CPU and pin the oracle.
Nothing in the product package imports this.
"""
from .models import create_model  # noqa: F401
