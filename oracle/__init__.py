"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU (fp32 / fp64) restatement of the MMNN_STS hot path used as the parity checker.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (mmnn_sts_b200/) never imports it.

Parity pin: the reference ships no tests or golden vectors, so the
pins are (a) outputs of the reference's OWN unchanged model / loss files executed in the
build container under oracle/shim.py, committed as tests/golden/*.npz together with the
generating script tests/golden/make_golden.py, and (b) hand-checked known-answer vectors
for the two third-party functions (pycox CoxPHLoss, lifelines concordance_index) whose
source is not vendored in the reference -- those two are restated from their published
algorithm ("parity pinned on KATs + reference call sites").
"""
