from . import reparameterization  # noqa: F401
