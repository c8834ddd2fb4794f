"""tensorflow_probability.python.internal.reparameterization: two marker constants.  TEST INFRASTRUCTURE ONLY."""
FULLY_REPARAMETERIZED = "FULLY_REPARAMETERIZED"
NOT_REPARAMETERIZED = "NOT_REPARAMETERIZED"
