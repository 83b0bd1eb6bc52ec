#!/bin/bash
mkdir -p gpurun_out
run() { # tag, args
  timeout 2400 python bench.py $2 > gpurun_out/r02i_$1.log 2> gpurun_out/r02i_$1.err; echo "$1 rc=$?"; tail -c 400 gpurun_out/r02i_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02i_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f e2e %.0f parity %s launches %d wave %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"], d["gpu_launches"], d["details"]["individuals_per_wave"]))
    print("  stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
    print("  peaks", json.dumps(d["details"]["measured_peaks"])[:400])
    if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:200], d["cpu_baseline"]["max_abs_fitness_diff_vs_gpu"])
    if d.get("parity"): print("  parity", json.dumps(d["parity"])[:500])
except Exception as e:
    print(tag, "no line", e)
PY
}
run c2 "--steps 5 --warmup 3 --no-cpu-baseline"
run c4p63 "--workload c4_20000x500000_k50000_pop500 --pop 63 --steps 2 --warmup 1 --no-cpu-baseline --no-parity"
run c4 "--workload c4_20000x500000_k50000_pop500 --steps 2 --warmup 1"
