#!/bin/bash
# final bench lines of round 2 after the last kernel changes (one GPU): C2 (driver-style), reference arm, C3, C4
mkdir -p gpurun_out
run() { # tag, timeout, args
  timeout $2 python bench.py $3 > gpurun_out/r02x_$1.log 2> gpurun_out/r02x_$1.err; echo "$1 rc=$?"; tail -c 300 gpurun_out/r02x_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02x_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.1f ms %.2f e2e %.1f parity %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("parity_ok"), d["gpu_launches"]))
    if "stage_ms_per_step" in d:
        print("  stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
        print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
    if d.get("cpu_baseline"): print("  cpu", round(d["cpu_baseline"]["value"], 4), d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("max_abs_fitness_diff_vs_gpu"))
    if d.get("parity"): print("  parity", d["parity"]["max_abs_fitness_diff_vs_exact_oracle"], d["parity"]["gram_bit_exact"], d["parity"]["genomes"])
except Exception as e:
    print(tag, "no line", e)
PY
}
run c2 600 "--steps 20 --warmup 5"
run ref 400 "--impl reference --steps 3 --warmup 1"
run c3 900 "--workload c3_5000x50000_k5001_pop1000_10fold --steps 3 --warmup 2"
run c4 1500 "--workload c4_20000x500000_k50000_pop500 --steps 2 --warmup 1"
