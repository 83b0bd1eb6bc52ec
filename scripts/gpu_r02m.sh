#!/bin/bash
mkdir -p gpurun_out
run() { # tag, args
  timeout 300 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-parity --no-sustained-peaks $2 > gpurun_out/r02m_$1.log 2> gpurun_out/r02m_$1.err; echo "$1 rc=$?"; tail -c 200 gpurun_out/r02m_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02m_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "ms %.2f" % d["ms_per_step"], "stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items() if k in ("gram", "solve", "chol_update", "chol_panel")})
except Exception as e:
    print(tag, "no line", e)
PY
}
run normal ""
run nostore "--opt gram_experiment=1"
run noepi "--opt gram_experiment=2"
