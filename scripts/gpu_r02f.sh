#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_packed.py -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r02f_tests.log
run() { # tag, args
  timeout 1200 python bench.py --steps 3 --warmup 2 $2 > gpurun_out/r02f_$1.log 2> gpurun_out/r02f_$1.err; echo "$1 rc=$?"; tail -c 300 gpurun_out/r02f_$1.err
  python - "$1" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r02f_%s.log" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f ms %.2f e2e %.0f parity %s launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"], d["gpu_launches"]))
    print("  stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    print("  fracs", {k: (round(d[k]["frac"], 3) if d[k]["frac"] else None) for k in d if k.startswith("roofline")})
    if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["max_abs_fitness_diff_vs_gpu"])
    if d.get("parity"): print("  parity", d["parity"]["max_abs_fitness_diff_vs_exact_oracle"], d["parity"]["gram_bit_exact"])
except Exception as e:
    print(tag, "no line", e)
PY
}
run c3 "--workload c3_5000x50000_k5001_pop1000_10fold --no-cpu-baseline"
timeout 900 python scripts/sweep.py --out gpurun_out/r02_sweep.json > gpurun_out/r02f_sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/r02f_sweep.log | cut -c1-300
