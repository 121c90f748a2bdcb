"""CPU oracle for the TensorBNN HMC hot path.  TEST INFRASTRUCTURE ONLY.

PARITY PINNED TO THE REFERENCE'S OWN SOURCE (round 2).  The reference (alpha-davidson/TensorBNN) ships no
tests, golden vectors or fixtures, and its arithmetic lives in TensorFlow / TensorFlow-Probability, which are
not installed here and cannot be (no network, no wheels).  Two layers of oracle therefore exist:

1. ``oracle/tfshim/`` -- torch-backed stand-ins for the ~70 ``tf.*`` / ``tfp.*`` / ``emcee`` symbols the
   reference touches.  Over them ``tests/golden/make_ref_golden.py`` imports /root/reference/tensorBNN/*.py
   UNMODIFIED and runs ``network.train`` -> ``stepMCMC`` (target closures, TFP-ordered leapfrog, MH, hyper
   chain, dual averaging), every layer / likelihood / density function, ``paramAdapter.update`` and the
   sample writer + ``predictor`` reader, committing the results as tests/golden/ref_*.json, reffn_*.json,
   ref_run/.  What remains a restatement is only TensorFlow's and TFP's own published semantics (the shim);
   the reference's code is executed, not paraphrased.
2. This package -- a restatement of the same algorithm (torch autograd in targets.py, hand-derived numpy
   gradients in analytic.py, TFP-HMC semantics in hmc.py, adapter.py, fileformat.py), each function citing the
   reference file:line it follows.  tests/test_ref_golden.py holds it to the reference-run fixtures (they agree
   to 0-1e-13 in fp64, incl. TF's float32-rounded python constants, tfconst.py); it is what the GPU parity
   tests and bench.py's cpu_baseline call at arbitrary sizes.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (tensorbnn_b200) never does; it fails loudly when its CUDA library is missing.
"""
