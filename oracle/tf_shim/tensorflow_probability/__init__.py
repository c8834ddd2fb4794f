"""Minimal stand-in for the `tensorflow_probability` symbols the reference's observation-model path uses.

TEST INFRASTRUCTURE ONLY -- see oracle/tf_shim/README.md.
"""
from . import distributions  # noqa: F401
from .distributions import push_uniforms, pending_uniforms  # noqa: F401

__version__ = "0.16-shim"
