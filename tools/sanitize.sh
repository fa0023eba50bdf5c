#!/bin/bash
# compute-sanitizer over the small fp32 / tensor-core / role-specialised / sparse loop tests (one tool per invocation):
#   tools/sanitize.sh memcheck|racecheck|synccheck > profiles/r2_sanitizer_<tool>.txt
# The persistent kernels spin on each other, and the sanitizer slows them 10-100x: the spin deadline is widened.
tool=${1:-memcheck}
export WRNN_SPIN_DEADLINE_MS=600000
exec compute-sanitizer --tool $tool --print-limit 20 python -m pytest -x -q \
    "tests/test_gpu_parity.py::test_mol_unbatched_vs_oracle" \
    "tests/test_gpu_tc.py::test_tc_loop_mol_teacher_forced_vs_reference" \
    "tests/test_gpu_rs.py::test_rs_loop_small_and_ragged_fold_counts" \
    "tests/test_gpu_sparse.py" 2>&1 | tail -40
