"""Constants as TensorFlow materialises them.  TEST INFRASTRUCTURE (see oracle/__init__.py).

``tf.cast(python_float, dtype)`` first converts the python number to a float32 tensor and only then casts
(tf.cast -> ops.convert_to_tensor(x) with no dtype), so every python constant the reference routes through
``tf.cast`` -- 2*pi and the 1e-8 clamp of multivariateLogProb (BNN_functions.py:23-24,30), the SquarePrelu /
Prelu hyper-prior constants (activationFunctions.py:144-145,301-306), FixedGaussianLikelihood's sd
(likelihood.py:161), the dual-averaging constants (network.py:241-248), the start step sizes (network.py:237) --
carries float32 rounding even when the network dtype is float64.  The same holds for the float32
``tfd.MultivariateNormalDiag(loc=[c], scale_diag=[s])`` hyper-priors of the dense layers (layer.py:137-153,
318-334; python lists -> float32).  Invisible in float32 (the default dtype); reproduced in float64 so that
"1e-10 relative in fp64" is meant against the reference's arithmetic, not against idealised constants
(quirk Q14).
"""
import math

import numpy as np


def f32(c):
    return float(np.float32(c))


TWO_PI_CAST = f32(2.0 * math.pi)          # tf.cast(2 * math.pi, dtype), BNN_functions.py:30
LOG_2PI_MVLP = math.log(TWO_PI_CAST)      # the k*log(2pi) term of multivariateLogProb
LOG_2PI = math.log(2.0 * math.pi)         # inside tfd.MultivariateNormalDiag.log_prob (exact in the dtype)
CLAMP_LO = f32(10 ** (-8))                # BNN_functions.py:23
CLAMP_HI = f32(10 ** 8)                   # BNN_functions.py:24 (exactly representable)
