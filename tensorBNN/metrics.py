"""Drop-in alias of tensorbnn_b200.metrics (same names as the reference module tensorBNN/metrics.py)."""
from tensorbnn_b200.metrics import *  # noqa: F401,F403
