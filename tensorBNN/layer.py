"""Drop-in alias of tensorbnn_b200.layer (same names as the reference module tensorBNN/layer.py)."""
from tensorbnn_b200.layer import *  # noqa: F401,F403
