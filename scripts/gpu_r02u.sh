#!/bin/bash
# round 2: L2 prefetch of the next tile's cross-products in the update epilogue (A/B)
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02u_tests.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/r02u_tests.log
for v in 1 0 1 0; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sustained-peaks --no-parity --opt upd_prefetch=$v > gpurun_out/r02u_pf_$v.log 2> gpurun_out/r02u_pf_$v.err; echo rc=$?
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02u_pf_$v.log") if l.startswith("{")][-1])
print("upd_prefetch=$v", d["value"], d["ms_per_step"], {k: round(x, 2) for k, x in d["stage_ms_per_step"].items()})
PY
done
