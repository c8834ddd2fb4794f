"""Placeholder for `from tensorflow.keras import layers` (models/model06.py:10).  TEST INFRASTRUCTURE ONLY."""


def __getattr__(name):
    raise RuntimeError("tf.keras.layers.%s is not part of oracle/tf_shim (networks are out of scope)" % name)
