#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_packed.py -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r02j_tests.log | cut -c1-300
run() { # tag, args
  timeout 1200 python bench.py $2 > gpurun_out/r02j_$1.log 2> gpurun_out/r02j_$1.err; echo "$1 rc=$?"; tail -c 300 gpurun_out/r02j_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02j_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f e2e %.0f parity %s launches %d wave %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"], d["gpu_launches"], d["details"]["individuals_per_wave"]))
    print("  stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
except Exception as e:
    print(tag, "no line", e)
PY
}
run c2 "--steps 5 --warmup 3 --no-cpu-baseline --no-sustained-peaks"
run c2single "--steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-sustained-peaks --opt gram_pair=0"
run c4p63 "--workload c4_20000x500000_k50000_pop500 --pop 63 --steps 2 --warmup 1 --no-cpu-baseline --no-sustained-peaks"
timeout 600 python scripts/knockout_bench.py --out gpurun_out/r02_knockout.json > gpurun_out/r02j_knockout.log 2>&1; echo "knockout rc=$?"; tail -2 gpurun_out/r02j_knockout.log | cut -c1-600
