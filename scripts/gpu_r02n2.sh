#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-sustained-peaks > gpurun_out/r02b_bench_n2.log 2> gpurun_out/r02b_bench_n2.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/r02b_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02b_bench_n2.log") if l.startswith("{")][-1])
    print("N=2 weak value %.0f ms %.2f e2e %.0f parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"]), d["details"]["cross_product_storage"])
    o = d["details"]["other_scaling"]
    print("strong: value %.0f ms %.2f e2e %.0f" % (o["value"], o["ms_per_step"], o["e2e_value"]), {k: round(v, 2) for k, v in o["stage_ms_per_step"].items()})
except Exception as e:
    print("no line", e)
PY
