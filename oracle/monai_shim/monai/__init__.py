"""Minimal stand-in for the handful of MONAI 0.5 names the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). It exists so the *unmodified*
reference files under /root/reference/source_code can be imported in a
container without MONAI (filters_and_operators.py:11-13,
stylization_layers.py:3-4). Semantics follow SURVEY.md Appendix B.
"""
__version__ = "0.5-shim"
