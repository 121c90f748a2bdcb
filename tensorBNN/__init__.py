"""Drop-in import path: `from tensorBNN.network import network` etc. resolve to tensorbnn_b200."""
