#!/bin/bash
# round 2, last kernels: full GPU test suite, then the profiling recipe at pop 1 000 (launch list + full captures)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02w_tests.log | cut -c1-200
bash scripts/profile.sh r02b mixed 1000 2>&1 | tail -20
