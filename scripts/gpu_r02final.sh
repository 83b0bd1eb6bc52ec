#!/bin/bash
# final measurement pass of round 2 (one GPU): full GPU tests, bench lines of every BASELINE config, sweep, knockout, main.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r02_final_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02_final_tests.log | cut -c1-200
run() { # tag, timeout, args
  timeout $2 python bench.py $3 > gpurun_out/r02_final_$1.log 2> gpurun_out/r02_final_$1.err; echo "$1 rc=$?"; tail -c 300 gpurun_out/r02_final_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02_final_%s.log" % tag) if l.startswith("{")][-1])
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
run c1 300 "--workload c1_1000x10000_k1500_pop50 --steps 20 --warmup 5"
run c3 900 "--workload c3_5000x50000_k5001_pop1000_10fold --steps 3 --warmup 2"
run c4 1500 "--workload c4_20000x500000_k50000_pop500 --steps 2 --warmup 1"
timeout 600 python scripts/sweep.py --out gpurun_out/r02_sweep.json > gpurun_out/r02_final_sweep.log 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/r02_final_sweep.log
timeout 300 python scripts/knockout_bench.py --out gpurun_out/r02_knockout.json > gpurun_out/r02_final_knockout.log 2>&1; echo "knockout rc=$?"; tail -1 gpurun_out/r02_final_knockout.log | cut -c1-500
timeout 900 python scripts/main_c1.py --json gpurun_out/r02_main_c1.json > gpurun_out/r02_final_main_c1.log 2>&1; echo "main_c1 rc=$?"; tail -2 gpurun_out/r02_final_main_c1.log | cut -c1-500
timeout 900 python scripts/main_c1.py --features 1200 --gens 3 --json gpurun_out/r02_main_c1_knockout.json --local_search knockout > gpurun_out/r02_final_main_ko.log 2>&1; echo "main_c1 knockout rc=$?"; tail -2 gpurun_out/r02_final_main_ko.log | cut -c1-500
