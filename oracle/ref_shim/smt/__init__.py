"""Minimal stand-in for the `smt` package so that the reference imports.

TEST INFRASTRUCTURE ONLY.  The reference imports ``smt.sampling_methods.LHS`` at
module scope (gpgradpy/src/optz/GpHparaX0.py:12) but `smt` is not installed in
this image.  Only start-point *selection* uses it; the hot path never does.
"""
