#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8.log 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc=$?"
tail -c 600 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_n8.log") if l.startswith("{")][-1])
    print("N=8 weak value %.0f ms %.2f e2e %.0f parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_ok"]), d["details"]["cross_product_storage"])
    print("stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    o = d["details"]["other_scaling"]
    print("strong: value %.0f ms %.2f e2e %.0f" % (o["value"], o["ms_per_step"], o["e2e_value"]), {k: round(v, 2) for k, v in o["stage_ms_per_step"].items()})
except Exception as e:
    print("no line", e)
PY
