"""`tf.keras` names that appear in class statements / module-level imports of models/model0N.py.  Nothing here
computes: the conv encoder / decoder are out of scope, the golden generator only calls `loss_fn` of models/model06.py.
TEST INFRASTRUCTURE ONLY."""
from . import layers  # noqa: F401


class Model:
    def __init__(self, *a, **k):
        raise RuntimeError("tf.keras.Model is a placeholder in oracle/tf_shim (networks are out of scope)")


class _Unavailable:
    def __getattr__(self, name):
        raise RuntimeError("tf.keras.%s is not part of oracle/tf_shim" % name)


optimizers = _Unavailable()
