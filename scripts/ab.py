"""A/B helper: median / min / max device time per step of the resident evaluation and the per-stage split.
usage: python scripts/ab.py [pop] [steps] [precision] [storage]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tblup_b200 import GblupEngine, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
prec = sys.argv[3] if len(sys.argv) > 3 else "mixed"
storage = sys.argv[4] if len(sys.argv) > 4 else "packed2"
x, y = synth.synth_dataset(5000, 50000, seed=0)
tr, va, te = synth.split_indices(5000, seed=0)
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]), storage=storage)
eng.set_rowset(0, tr, va)
eng.set_precision(prec)
for kv in os.environ.get("TB_OPTS", "").split(","):      # e.g. TB_OPTS=fuse_scale=0
    if "=" in kv:
        eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
stream = torch.cuda.current_stream()
eng.set_stream(stream.cuda_stream)
flat, off = synth.random_genomes(P, 50000, 5001, seed=1)
eng.stage(flat=flat, off=off)
fit = torch.empty(P, dtype=torch.float64, device="cuda")
for _ in range(3):
    eng.evaluate_staged([0], out_device_ptr=fit.data_ptr())
ts = []
issue = []
for i in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.evaluate_staged([0], out_device_ptr=fit.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
    issue.append(eng.info("last_issue_us") / 1e3)
ts = np.array(ts)
print("pop %d %s %s: per-step ms median %.1f min %.1f max %.1f  -> %.0f evals/s (median)" % (
    P, prec, storage, np.median(ts), ts.min(), ts.max(), P / np.median(ts) * 1e3))
eng.set_option("profile", 1)
agg = {}
for i in range(5):
    eng.reset_counters()
    eng.evaluate_staged([0], out_device_ptr=fit.data_ptr())
    for k, v in eng.stage_times().items():
        agg.setdefault(k, []).append(v[0])
print("per-step ms:", [round(float(t), 1) for t in ts], "fallbacks", eng.info("last_fallbacks"))
print("host issue ms per step:", [round(t, 2) for t in issue])
print("stage medians (ms):", {k: round(float(np.median(v)), 2) for k, v in agg.items()},
      "sum %.1f" % sum(float(np.median(v)) for v in agg.values()))
