"""Drop-in alias of tensorbnn_b200.likelihood (same names as the reference module tensorBNN/likelihood.py)."""
from tensorbnn_b200.likelihood import *  # noqa: F401,F403
