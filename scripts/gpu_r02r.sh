#!/bin/bash
# round 2, t16: block-column entries below the diagonal block as halves (A/B against t16=0)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02r_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02r_tests.log
for v in 1 0; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sustained-peaks --opt t16=$v > gpurun_out/r02r_t16_$v.log 2> gpurun_out/r02r_t16_$v.err; echo rc=$?
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02r_t16_$v.log") if l.startswith("{")][-1])
print("t16=$v", d["value"], d["ms_per_step"], d["parity_ok"], d["details"].get("stage_ms_per_step"), d.get("parity", {}).get("max_abs_diff_vs_exact_oracle") if isinstance(d.get("parity"), dict) else None)
PY
done
