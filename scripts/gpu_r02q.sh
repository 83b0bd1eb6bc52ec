#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sustained-peaks > gpurun_out/r02q_c2.log 2> gpurun_out/r02q_c2.err; echo rc=$?; tail -c 300 gpurun_out/r02q_c2.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02q_c2.log") if l.startswith("{")][-1])
print(d["value"], d["parity_ok"], d["details"]["device_de"])
PY
timeout 600 python scripts/sweep.py --pops 13,125,500 --out gpurun_out/r02_sweep_pergpu_of8.json 2>&1 | tail -1
