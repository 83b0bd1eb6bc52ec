#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k "two_ctas or small_matrices or fitness_matches" > gpurun_out/r02o_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/r02o_smoke.log | cut -c1-200
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 python -m pytest tests/test_gpu_large.py -m gpu -x -q -k "config4" > gpurun_out/r02o_c4test.log 2>&1; echo "c4 test rc=$?"; tail -3 gpurun_out/r02o_c4test.log | cut -c1-200
run() { # tag, args
  timeout 500 python bench.py $2 > gpurun_out/r02o_$1.log 2> gpurun_out/r02o_$1.err; echo "$1 rc=$?"; tail -c 200 gpurun_out/r02o_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02o_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f parity %s" % (d["value"], d["ms_per_step"], d["parity_ok"]), {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
except Exception as e:
    print(tag, "no line", e)
PY
}
run c4p63 "--workload c4_20000x500000_k50000_pop500 --pop 63 --steps 2 --warmup 1 --no-cpu-baseline --no-sustained-peaks"
