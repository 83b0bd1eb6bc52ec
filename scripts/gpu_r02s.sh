#!/bin/bash
# round 2: 16 epilogue warps + trimmed per-entry work in the Cholesky GEMM epilogue (A/B against epi_warps=8)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02s_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02s_tests.log
for v in 16 8; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sustained-peaks --opt epi_warps=$v > gpurun_out/r02s_ew_$v.log 2> gpurun_out/r02s_ew_$v.err; echo rc=$?
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02s_ew_$v.log") if l.startswith("{")][-1])
print("epi_warps=$v", d["value"], d["ms_per_step"], d["parity_ok"], d["parity"]["max_abs_fitness_diff_vs_exact_oracle"], {k: round(x, 2) for k, x in d["stage_ms_per_step"].items()})
PY
done
