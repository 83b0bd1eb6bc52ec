#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "paired" > gpurun_out/r02h_pair.log 2>&1; echo "pair tests rc=$?"
tail -12 gpurun_out/r02h_pair.log | cut -c1-400
run() { # tag, args
  timeout 600 python bench.py --steps 5 --warmup 3 $2 > gpurun_out/r02h_$1.log 2> gpurun_out/r02h_$1.err; echo "$1 rc=$?"; tail -c 300 gpurun_out/r02h_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02h_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f e2e %.0f parity %s launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"], d["gpu_launches"]))
    print("  stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
except Exception as e:
    print(tag, "no line", e)
PY
}
run cg2 "--no-cpu-baseline --opt gram_pair=2"
run single "--no-cpu-baseline --no-parity --opt gram_pair=0"
