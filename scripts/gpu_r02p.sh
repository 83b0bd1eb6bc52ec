#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "two_ctas or small_matrices or fitness_matches or folds" > gpurun_out/r02p_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r02p_smoke.log | cut -c1-200
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02p_tests.log | cut -c1-200
run() { # tag, args
  timeout 500 python bench.py $2 > gpurun_out/r02p_$1.log 2> gpurun_out/r02p_$1.err; echo "$1 rc=$?"; tail -c 200 gpurun_out/r02p_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02p_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f parity %s launches %d" % (d["value"], d["ms_per_step"], d["parity_ok"], d["gpu_launches"]), {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(tag, "no line", e)
PY
}
run k0 "--steps 8 --warmup 3 --no-cpu-baseline --no-sustained-peaks"
run scale32 "--steps 8 --warmup 3 --no-cpu-baseline --no-parity --no-sustained-peaks --opt blk0_scale32=1"
run c3 "--workload c3_5000x50000_k5001_pop1000_10fold --steps 2 --warmup 1 --no-cpu-baseline --no-sustained-peaks"
timeout 600 python scripts/startup_bench.py --out gpurun_out/r02_startup.json 2>&1 | tail -1 | cut -c1-700
