"""CPU oracle for the TensorBNN HMC hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (alpha-davidson/TensorBNN) ships no tests, no
golden vectors and no fixtures, and its arithmetic lives in TensorFlow /
TensorFlow-Probability, which are not installed here and cannot be (no
network, no wheels).  This package is therefore a *restatement* of the
reference's algorithm, each function citing the reference file:line it
follows; TFP's HMC semantics are restated from its published algorithm
(SURVEY.md Appendix B).  It is pinned only against hand-derived closed forms,
central finite differences and a second, independently written analytic
gradient (oracle/analytic.py) -- see tests/test_oracle_*.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (tensorbnn_b200) never
does; it fails loudly when its CUDA library is missing.
"""
