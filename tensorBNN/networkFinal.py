"""The reference Examples import tensorBNN.networkFinal (a module missing from its tree)."""
from tensorbnn_b200.network import network  # noqa: F401
