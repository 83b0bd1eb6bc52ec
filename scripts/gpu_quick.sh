#!/bin/bash
# quick GPU check: parity tests (optionally a subset via $TESTS), then one bench line without the CPU arm
TAG=${TAG:-quick}
mkdir -p gpurun_out
timeout 1500 python -m pytest ${TESTS:-tests} -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json, os
tag = os.environ.get("TAG", "quick")
try:
    d = json.loads([l for l in open("gpurun_out/%s_bench.log" % tag) if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity_ok"])
    print("stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    for k in d:
        if k.startswith("roofline"):
            print(k, d[k]["kernel"][:28], "frac", d[k]["frac"], "sweeps", d[k].get("mean_refinement_sweeps"))
except Exception as e:
    print("no bench line:", e)
PY
