#!/bin/bash
# round 2: diag32 with 128 threads / 7 blocks per SM; epilogue warps on the int32 cross-product path
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02t_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02t_tests.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sustained-peaks > gpurun_out/r02t_c2.log 2> gpurun_out/r02t_c2.err; echo rc=$?
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02t_c2.log") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["parity_ok"], d["parity"]["max_abs_fitness_diff_vs_exact_oracle"], d["parity"]["fp64_fallbacks_in_sample"], {k: round(x, 2) for k, x in d["stage_ms_per_step"].items()})
PY
for v in 16 8; do
timeout 300 python scripts/sweep.py --ks 10000 --pops 125,500 --opt epi_warps=$v --out gpurun_out/r02t_sweep_ew$v.json 2>&1 | grep -v "^crossover" | cut -c1-600
done
