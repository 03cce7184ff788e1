"""Diagnostic: kernel timeline (CUPTI) of two pipelined device-collated KvsAll steps -> gpurun_out/timeline_dc.txt."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import kgeb200 as kb
from torch.profiler import profile, ProfilerActivity

B = 4096
graph, batches = bench.build_batches(10, B, seed=7)
E, R = graph["num_entities"], graph["num_relations"]
nnz_max = max(int(b["label_coords"].shape[0]) for b in batches)
dev = torch.device("cuda")
model = kb.KgeModel("complex", E, R, 128).to(dev)
opt = kb.optim.create("Adagrad", model.parameters(), lr=0.2)
job = kb.TrainingJobKvsAll(model, opt, kb.KgeLoss.create("bce"), E, R, fused_path=True, math_mode=kb.lib.MATH_BF16)
job.enable_graph_step(B, nnz_max)
job.enable_device_collate(*graph["_indexes"])
ids = [b["example_ids"].pin_memory() for b in batches]
job.prefetch_ids(ids[0])
for i in range(5):
    job.prefetch_ids(ids[i + 1]); job.step_ids()
torch.cuda.synchronize()
import time
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter()
    for i in range(5, 8):
        job.prefetch_ids(ids[i + 1]); job.step_ids()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 3
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
with open(os.path.join(ROOT, "gpurun_out", "timeline_dc.txt"), "w") as f:
    f.write(f"# wall per step {wall * 1e6:.1f} us (3 steps, no sync between them)\n")
    for e in evs:
        st = e.time_range.start - t0
        f.write(f"{st:9.1f} {e.device_time:8.1f} {st + e.device_time:9.1f}  {e.name[:100]}\n")
print("wall per step us", wall * 1e6)
