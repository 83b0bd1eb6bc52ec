#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "paired or fp4_gram" > gpurun_out/r02g_pair.log 2>&1; echo "pair tests rc=$?"
tail -12 gpurun_out/r02g_pair.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_packed.py -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02g_tests.log | cut -c1-300
run() { # tag, args
  timeout 600 python bench.py --steps 5 --warmup 3 $2 > gpurun_out/r02g_$1.log 2> gpurun_out/r02g_$1.err; echo "$1 rc=$?"; tail -c 300 gpurun_out/r02g_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02g_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f e2e %.0f parity %s launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"], d["gpu_launches"]))
    print("  stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
except Exception as e:
    print(tag, "no line", e)
PY
}
run pair "--no-cpu-baseline"
run nopair "--no-cpu-baseline --no-parity --opt gram_pair=0"
