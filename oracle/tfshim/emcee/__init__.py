"""Stand-in for ``emcee`` (only ``emcee.autocorr`` is used, predictor.py:6).  TEST INFRASTRUCTURE ONLY."""
from . import autocorr  # noqa: F401
