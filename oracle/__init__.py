"""CPU oracle for the WaveRNN vocoder inference hot path.

THIS DIRECTORY IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import or execute anything below oracle/ -- and only as the
checker, never as the thing shipped.  The product package (real-time-voice-cloning_b200/) never
imports it and fails loudly when its CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors ("parity unpinned by the
reference's own tests").  The restatement in wavernn_oracle.py is therefore pinned against outputs of
the UNMODIFIED reference imported from /root/reference in the build container
(oracle/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
"""
