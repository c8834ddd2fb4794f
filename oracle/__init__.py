"""CPU oracle for the vae-mdl observation-model hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only
as the checker or as the timed CPU baseline, never as something the product
path routes through.  The product (``vae_mdl_b200``) raises when its CUDA
library is missing; it never falls back to this package.

PARITY UNPINNED: the reference (nbip/vae-mdl) ships no golden vectors,
known-answer tests or fixtures for this path (its ``tests/`` directory holds
assert-free ``__main__`` scripts), and TensorFlow / TensorFlow-Probability are
not installable in this image, so the reference itself cannot be executed to
generate fixtures.  The oracle is therefore an op-for-op restatement of the
reference's formulas (each function cites the reference file:line it follows),
cross-validated three ways (see ``tests/test_oracle.py``): the two independent
formulations in the reference (``utils/mdl.py`` vs ``utils/mdl_openai.py``)
agree with each other, the float64 flavour agrees with an mpmath
arbitrary-precision evaluation of the closed form on hand-built branch cases,
and analytic gradients agree with autograd.
"""
from .ref import *  # noqa: F401,F403
