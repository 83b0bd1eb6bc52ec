"""Per-step device times of the resident evaluation (diagnostic: run-to-run variation, clock behaviour)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tblup_b200 import GblupEngine, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
x, y = synth.synth_dataset(5000, 50000, seed=0)
tr, va, te = synth.split_indices(5000, seed=0)
eng = GblupEngine(x, y, perm=np.concatenate([tr, va, te]))
eng.set_rowset(0, tr, va)
stream = torch.cuda.current_stream()
eng.set_stream(stream.cuda_stream)
flat, off = synth.random_genomes(P, 50000, 5001, seed=1)
eng.stage(flat=flat, off=off)
fit = torch.empty(P, dtype=torch.float64, device="cuda")
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
for i in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    eng.evaluate_staged([0], out_device_ptr=fit.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print("step %2d  device %.1f ms  wall %.1f ms  sm %d MHz  power %.0f W" % (
        i, e0.elapsed_time(e1), wall, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
        pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
