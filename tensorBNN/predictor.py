"""Drop-in alias of tensorbnn_b200.predictor (same names as the reference module tensorBNN/predictor.py)."""
from tensorbnn_b200.predictor import *  # noqa: F401,F403
