"""Minimal stand-in for the MONAI symbols the reference's hot-path modules import.

Test infrastructure ONLY: MONAI is not installed in this image and there is no network, so the unmodified reference
under /root/reference is imported on top of this shim to (a) validate oracle/torch_oracle.py and (b) generate the
golden vectors under tests/golden/.  Each class follows the documented MONAI semantics listed in SURVEY.md §8(c);
nothing here is used by the product package.
"""
__version__ = "0.0-shim"
