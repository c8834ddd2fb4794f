"""CPU oracle for the vae-mdl observation-model hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: it may be
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only
as the checker or as the timed CPU baseline, never as something the product
path routes through.  The product (``vae_mdl_b200``) raises when its CUDA
library is missing; it never falls back to this package.

PARITY PINNED TO THE REFERENCE'S SOURCE, NOT TO TENSORFLOW'S KERNELS.  The
reference (nbip/vae-mdl) ships no golden vectors, known-answer tests or fixtures
for this path (its ``tests/`` directory holds assert-free ``__main__`` scripts),
and TensorFlow / TensorFlow-Probability are not installable in this image.  What
pins the oracle instead:

* ``tests/golden/refsrc_*.npz`` -- outputs of the reference's OWN, unmodified
  modules (``utils/mdl.py``, ``mdl_openai.py``, ``mdl_openai_iwae.py``,
  ``discretized_logistic.py``, ``mdl_plain.py``, ``utils.py``,
  ``models/loss.py``, ``models/model06.py::loss_fn``) imported from
  ``/root/reference`` and executed over ``oracle/tf_shim`` -- a torch-CPU
  stand-in for the ~45 ``tf`` / ``tfd`` primitives those modules call
  (``tests/golden/make_reference_golden.py``; float64 and float32 runs,
  gradients by autograd over the executed code).  The oracle agrees with them to
  round-off (``tests/test_reference_golden.py``) and the CUDA path is checked
  against the same vectors.  What stays unpinned is TensorFlow's float32
  rounding inside its own ``sigmoid`` / ``exp`` / ``log`` kernels (an ulp or
  two, far inside the 1e-5 / 1e-4 tolerances).
* ``tests/test_oracle.py``: the two independent formulations inside the
  reference (``utils/mdl.py`` vs ``utils/mdl_openai.py``) agree with each other,
  the float64 flavour agrees with an mpmath arbitrary-precision evaluation of
  the closed form on hand-built branch cases, and analytic gradients agree with
  autograd / finite differences.
"""
from .ref import *  # noqa: F401,F403
