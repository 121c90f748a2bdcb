"""Drop-in alias of tensorbnn_b200.network (same names as the reference module tensorBNN/network.py)."""
from tensorbnn_b200.network import *  # noqa: F401,F403
