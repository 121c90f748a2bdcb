"""Drop-in alias of tensorbnn_b200.activationFunctions (same names as the reference module tensorBNN/activationFunctions.py)."""
from tensorbnn_b200.activationFunctions import *  # noqa: F401,F403
