#!/bin/bash
# guarded run: a short smoke of the two-CTA solve first (strict timeout), everything else only if it passes
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "two_ctas or fitness_matches_reference or small_matrices" > gpurun_out/r02l_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"
tail -5 gpurun_out/r02l_smoke.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "smoke failed: stopping"; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_packed.py tests/test_gpu_knockout.py tests/test_gpu_large.py -m gpu -x -q > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r02l_tests.log | cut -c1-400
run() { # tag, args
  timeout 400 python bench.py $2 > gpurun_out/r02l_$1.log 2> gpurun_out/r02l_$1.err; echo "$1 rc=$?"; tail -c 300 gpurun_out/r02l_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02l_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f e2e %.0f parity %s launches %d wave %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"], d["gpu_launches"], d["details"]["individuals_per_wave"]))
    print("  stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
except Exception as e:
    print(tag, "no line", e)
PY
}
run p125pair "--steps 5 --warmup 3 --pop 125 --no-cpu-baseline --no-sustained-peaks"
run p125single "--steps 5 --warmup 3 --pop 125 --no-cpu-baseline --no-parity --no-sustained-peaks --opt solve_pair=0"
run c4p63 "--workload c4_20000x500000_k50000_pop500 --pop 63 --steps 2 --warmup 1 --no-cpu-baseline --no-sustained-peaks"
