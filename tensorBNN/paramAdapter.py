"""Drop-in alias of tensorbnn_b200.paramAdapter (same names as the reference module tensorBNN/paramAdapter.py)."""
from tensorbnn_b200.paramAdapter import *  # noqa: F401,F403
