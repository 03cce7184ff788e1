#!/usr/bin/env python
"""The FB15k-237-shaped workload (BASELINE.json configs[1]); round 1's bench line, now reported by bench.py under the
`extra` key (bench.py's primary workload is the one the north star scales over GPUs, configs[3]).  `python bench_fb237.py`
still prints its own line; `run(args)` returns it as a dict.

Benchmark of the hot path on the workload BASELINE.json quotes for one GPU (configs[1]):
ComplEx KvsAll + BCE, dim 128, synthetic FB15k-237-shaped graph (14,541 entities, 237 relations, 272,115
train triples).  A step = one KvsAll batch (gather -> query transform -> fused score+loss -> backward ->
scatter -> Adagrad on both tables).  Metric: training queries/s (a KvsAll example = one sp_ / _po query).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--math bf16|tf32|fp32]

Prints ONE JSON line (see the build contract): `value` = device-resident inputs, CUDA-event timed;
`e2e` = through TrainingJobKvsAll.step() with pinned host batches, host->device copies and the loss read
back inside the timed region; `roofline` for the dominant kernel; `cpu_baseline` = the oracle (a torch-CPU
port of the reference's path) on a bounded sample on this box's host cores.
`--impl reference` times that CPU port alone (the reference is pure Python and does not travel to the box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "ComplEx KvsAll+BCE d=128, synthetic FB15k-237 shape (E=14541, R=237, 272115 train triples)"
DIM, LR = 128, 0.2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------
# workload: graph, KvsAll examples, batches (host side, vectorised; the reference does this in Python loops)
# ---------------------------------------------------------------------------------------------
def build_batches(num_batches, batch_size, seed, rank=0):
    import kgeb200 as kb
    g = kb.graph.synthetic_graph("fb15k-237", seed=0)
    idx = [kb.index.KvsAllIndex(g["train"], "sp"), kb.index.KvsAllIndex(g["train"], "po")]
    n_sp = len(idx[0])
    n_ex = n_sp + len(idx[1])
    rng = np.random.default_rng(seed + 1000 * rank)
    batches = []
    keys = [i._keys.numpy() for i in idx]
    offs = [i._values_offset.numpy() for i in idx]
    vals = [i._values.numpy() for i in idx]
    for _ in range(num_batches):
        ex = rng.choice(n_ex, batch_size, replace=False)
        qt = (ex >= n_sp).astype(np.int64)
        loc = np.where(qt == 0, ex, ex - n_sp)
        queries = np.where(qt[:, None] == 0, keys[0][np.minimum(loc, len(keys[0]) - 1)],
                           keys[1][np.minimum(loc, len(keys[1]) - 1)])
        start = np.where(qt == 0, offs[0][np.minimum(loc, len(keys[0]) - 1)], offs[1][np.minimum(loc, len(keys[1]) - 1)])
        end = np.where(qt == 0, offs[0][np.minimum(loc, len(keys[0]) - 1) + 1],
                       offs[1][np.minimum(loc, len(keys[1]) - 1) + 1])
        n = end - start
        lab_off = np.zeros(batch_size + 1, dtype=np.int64)
        lab_off[1:] = np.cumsum(n)
        rows = np.repeat(np.arange(batch_size), n)
        pos = np.arange(lab_off[-1]) - lab_off[rows] + start[rows]
        lab = np.where(qt[rows] == 0, vals[0][np.minimum(pos, len(vals[0]) - 1)], vals[1][np.minimum(pos, len(vals[1]) - 1)])
        coords = np.stack([rows, lab], 1).astype(np.int32)
        batches.append({"example_ids": torch.from_numpy(ex.astype(np.int64)),
                        "queries": torch.from_numpy(queries.astype(np.int64)),
                        "label_coords": torch.from_numpy(coords),
                        "query_type_indexes": torch.from_numpy(qt)})
    g["_indexes"] = idx
    return g, batches


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines, self.index, self.n0 = None, [], index, 0

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def __enter__(self):
        self.t_enter = time.time()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
            # nvidia-smi needs a few hundred ms before its first line: wait for it HERE (before the timed region), so that a
            # region of ~0.1 s is sampled too; lines from before the region are not used unless nothing else arrived
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:
                time.sleep(0.01)
            self.n0 = len(self.lines)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in (self.lines[self.n0:] or self.lines):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx.append(int(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (torch-CPU port of the reference's KvsAll batch body + Adagrad)
# ---------------------------------------------------------------------------------------------
def cpu_reference(batches, graph, steps, warmup, batch_size):
    from oracle import kge_oracle as ko
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    e, r = graph["num_entities"], graph["num_relations"]
    gen = torch.Generator().manual_seed(0)
    prm = ko.Params(torch.randn(e, DIM, generator=gen) * 0.1, torch.randn(r, DIM, generator=gen) * 0.1)
    opt = ko.make_optimizer("Adagrad", prm, lr=LR)
    times = []
    for i in range(warmup + steps):
        b = batches[i % len(batches)]
        t0 = time.perf_counter()
        opt.zero_grad()
        ko.batch_kvsall("complex", prm, b["queries"].numpy(), b["label_coords"].numpy(), b["query_type_indexes"].numpy(),
                        e, r, "bce")
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return {"value": batch_size / (ms / 1e3), "ms_per_step": ms, "cores": cores}


# ---------------------------------------------------------------------------------------------
def parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="KvsAll queries per step per GPU")
    ap.add_argument("--math", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--cpu-steps", type=int, default=8, help="steps of the bounded CPU-baseline sample")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="tuning runs: device-resident value + kernel roofline only")
    ap.add_argument("--parallel", default="p2p", choices=["dp", "p2p", "shard"],
                    help="N>1: dp = data-parallel replicas + one NCCL all-reduce of the gradients (graphs too small to "
                         "shard); p2p = the same replicas with the exchange fused into the Adagrad update over NVLink "
                         "peer memory, whole step in one CUDA graph; shard = entity-sharded scoring (SURVEY.md 8e)")
    ap.add_argument("--profile-calls", action="store_true", help="print GPU time per C-ABI call of one step and exit")
    ap.add_argument("--timeline", action="store_true",
                    help="write the kernel timeline of one step (CUPTI) to gpurun_out/timeline_bench.txt and exit")
    return ap


def main():
    out = run(parser().parse_args())
    if out is not None:
        print(json.dumps(out))


def run(args):
    """Runs the benchmark; returns the JSON line as a dict on rank 0 (None elsewhere / for the diagnostic modes)."""
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = min(args.steps, 10), min(args.warmup, 2)
        graph, batches = build_batches(steps + warm, B, seed=7)
        res = cpu_reference(batches, graph, steps, warm, B)
        sample = f"{steps} KvsAll steps of {B} queries (torch-CPU port of the reference's path; oracle/kge_oracle.py)"
        return ({
            "impl": "reference", "metric": "training queries/s", "value": res["value"], "unit": "queries/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch": B, "optimizer": "Adagrad lr 0.2"},
            "cpu_baseline": {"value": res["value"], "unit": "queries/s", "cores": res["cores"], "kind": "port",
                             "sample": sample},
            "e2e": {"value": res["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})

    import kgeb200 as kb
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    math_mode = {"bf16": kb.lib.MATH_BF16, "tf32": kb.lib.MATH_TF32, "fp32": kb.lib.MATH_FP32}[args.math]
    n_batches = args.steps + args.warmup
    # entity-sharded scoring: every rank processes the whole global batch (world * B queries) against its own
    # entity range, so all ranks build the same batches (weak scaling: per-GPU work B x E stays fixed)
    sharded = world > 1 and args.parallel == "shard"
    GB = B * world if sharded else B          # rows each rank processes per step
    graph, batches = build_batches(n_batches, GB, seed=7, rank=0 if sharded else rank)
    E, R = graph["num_entities"], graph["num_relations"]
    nnz_max = max(int(b["label_coords"].shape[0]) for b in batches)
    for b in batches:
        for k in b:
            b[k] = b[k].pin_memory()

    torch.manual_seed(0)
    model = kb.KgeModel("complex", E, R, DIM).to(dev)
    opt = kb.optim.create("Adagrad", model.parameters(), lr=LR)
    shard = kb.fused.Shard.of_rank(E, rank, world, dist.group.WORLD) if sharded else None
    job = kb.TrainingJobKvsAll(model, opt, kb.KgeLoss.create("bce"), E, R, fused_path=True, math_mode=math_mode,
                               shard=shard)
    try:
        job.enable_graph_step(GB, nnz_max, use_graph=not args.no_graph,
                              dp_group=dist.group.WORLD if (world > 1 and not sharded) else None,
                              dp_p2p=args.parallel == "p2p")
    except (RuntimeError, ImportError) as exc:
        # peer-mapped (symmetric) memory is a property of the box (NVLink / NVSwitch + fabric handles): where it cannot be
        # set up -- on every rank alike -- the replicas exchange gradients through NCCL instead, and the line says so
        if not (world > 1 and args.parallel == "p2p"):
            raise
        print(f"[bench] peer-memory exchange unavailable ({type(exc).__name__}: {exc}); using the NCCL mode", file=sys.stderr)
        args.parallel = "dp"
        job.enable_graph_step(GB, nnz_max, use_graph=not args.no_graph, dp_group=dist.group.WORLD, dp_p2p=False)
    stepper = job.stepper

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    if args.profile_calls:
        profile_calls(kb, job, stepper, batches)
        return
    if args.timeline:
        step_timeline(job, stepper, batches, flush)
        return

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM, CUDA-event timed -----------------------------
    dev_inputs = [job.device_inputs(b) for b in batches]
    for i in range(args.warmup):
        stepper.set_inputs(*dev_inputs[i])
        stepper.step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for i in range(args.steps):
            stepper.set_inputs(*dev_inputs[args.warmup + i])
            flush.fill_(i & 0xFF)           # evict L2 between timed steps (not timed)
            evs[i][0].record()
            stepper.step()
            evs[i][1].record()
        barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(np.sum(step_ms))
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = t.item()
    value = world * B * args.steps / (total_ms / 1e3)
    final_loss = (stepper.loss if getattr(stepper, "p2p", None) is None else stepper.loss_global).item()
    if getattr(stepper, "p2p", None) is not None:
        stepper.check_p2p()

    # ---------------- e2e: public API, pinned host batches, H2D + loss D2H inside the timed region -------
    e2e_s = 0.0
    packed = [job.collate_packed(b) for b in batches]   # host collate output (pinned), as a DataLoader worker emits it
    for i in range(0 if args.skip_e2e else args.warmup + args.steps):
        b = packed[i]
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # (job.prefetch_packed() -- the next batch's H2D copy underneath the running step -- measured no better here:
        # 21.5 M against 24.3 M queries/s; with a synchronize per step its D2D copy and events cost more than they hide)
        res = job.step(i, b)            # copies the batch to the device, runs the step, reads the loss back
        torch.cuda.synchronize()
        if i >= args.warmup:
            e2e_s += time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = None if args.skip_e2e else world * B * args.steps / e2e_s
    h2d = int(packed[0]["packed"].numel())

    # ---------------- e2e with on-device batch construction: the host sends example ids only ------------------
    #  batch i+1 is built on the collate stream while step i runs; every timed step contains one
    # H2D copy of 8*B bytes of ids, one batch construction, one step and the loss read-back
    e2e_dc = None
    if world == 1 and not args.skip_e2e:
        job.enable_device_collate(*graph["_indexes"])
        ids = [b["example_ids"].pin_memory() for b in batches]
        dc_s = 0.0
        job.prefetch_ids(ids[0])
        for i in range(args.warmup + args.steps):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if i + 1 < len(ids):
                job.prefetch_ids(ids[i + 1])
            job.step_ids()
            torch.cuda.synchronize()
            if i >= args.warmup:
                dc_s += time.perf_counter() - t0
        e2e_dc = {"value": world * B * args.steps / dc_s, "unit": "queries/s", "h2d_bytes_per_step": 8 * B,
                  "d2h_bytes_per_step": 8, "note": "KvsAll batches built on the device from example ids "
                  "(kgeb_kvsall_batch_*), double-buffered on a collate stream"}

    # ---------------- roofline of the dominant kernel (timed alone with CUDA events) ---------------------
    roof = kernel_roofline(kb, stepper, math_mode, GB, E)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- cpu baseline: bounded sample of the same workload on the host cores -----------------
    cpu = cpu_reference(batches, graph, args.cpu_steps, 1, B) if args.cpu_steps > 0 else {"value": None, "cores": 0}
    out = {
        "metric": "training queries/s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.math], "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "optimizer": "Adagrad lr 0.2",
                   "math": args.math + " tensor tiles, fp32 accumulate, fp32 master tables and optimizer",
                   "parallelism": "1 GPU" if world == 1 else (
                       f"entity-sharded scoring over {world} GPUs (all-reduce of row statistics, dQ, dense gradient)"
                       if sharded else (f"dp{world}: replicas; gradient exchange fused with the Adagrad update over "
                                        "NVLink peer memory (reduce-scatter of gradients, all-gather of updated "
                                        "weights, sharded optimizer state), one CUDA graph per step"
                                        if args.parallel == "p2p" else
                                        f"dp{world}: replicas with one all-reduce of both tables' gradients + loss per "
                                        "step (FB15k-237-sized tables are too small to shard, SURVEY.md 8e)")),
                   "l2": "flushed between timed steps (256 MiB write, untimed); table is 7.4 MB",
                   "cuda_graph": stepper.graph is not None, "final_loss": final_loss},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "note": "TrainingJobKvsAll.step() on packed pinned host batches; each timed step contains the H2D copy of "
                        "its batch, the step and the loss read-back"},
        "e2e_device_collate": e2e_dc,
        "gpu_launches": int(stepper.kernel_launches_per_step * args.steps),
        "roofline": roof,
        "cpu_baseline": {"value": cpu["value"], "unit": "queries/s", "cores": cpu["cores"], "kind": "port",
                         "sample": f"{args.cpu_steps} KvsAll steps of {B} queries, same graph/batches, torch-CPU port "
                                   "of the reference's path (oracle/kge_oracle.py)"},
    }
    if world > 1 and not getattr(args, "keep_process_group", False):
        dist.destroy_process_group()
    return out


def profile_calls(kb, job, stepper, batches):
    """GPU time of every C-ABI call of a step (sync + CUDA events around each call; eager mode)."""
    import collections
    acc = collections.OrderedDict()
    orig = kb.lib.call

    def timed(name, *a):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name, *a); e1.record()
        torch.cuda.synchronize()
        key = name + ("" if name != "kgeb_fused_bwd" else ("(dQ+dTable)"))
        acc.setdefault(key, []).append(e0.elapsed_time(e1) * 1e3)

    for m in (kb.lib, kb.trainer.lib):
        m.call = timed
    for i in range(6):
        stepper.set_inputs(*job.device_inputs(batches[i]))
        stepper._launch()
        if i == 1:
            acc.clear()
    for m in (kb.lib, kb.trainer.lib):
        m.call = orig
    tot = 0.0
    for k, v in acc.items():
        per_step = float(np.sum(v)) / 4
        tot += per_step
        print(f"{per_step:9.1f} us/step  {len(v) // 4} call(s)  {k}")
    print(f"{tot:9.1f} us/step  total of C-ABI calls")


def step_timeline(job, stepper, batches, flush):
    """Start / duration / stream of every kernel of one step as CUPTI sees it (diagnostic; not a bench number)."""
    from torch.profiler import profile, ProfilerActivity
    for i in range(3):
        stepper.set_inputs(*job.device_inputs(batches[i]))
        stepper.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3, 6):
            stepper.set_inputs(*job.device_inputs(batches[i]))
            flush.fill_(i)
            torch.cuda.synchronize()
            stepper.step()
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # last step = everything after the last L2-flush fill
    last = max(i for i, e in enumerate(evs) if "FillFunctor<unsigned char>" in e.name)
    evs = evs[last + 1:]
    t0 = evs[0].time_range.start
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rank = int(os.environ.get("RANK", "0"))
    with open(os.path.join(ROOT, "gpurun_out", "timeline_bench.txt" if rank == 0 else f"timeline_bench_rank{rank}.txt"), "w") as f:
        f.write("# start_us  dur_us  end_us  name   (one CUDA-graph replay of the training step)\n")
        for e in evs:
            st = e.time_range.start - t0
            f.write(f"{st:9.1f} {e.device_time:8.1f} {st + e.device_time:9.1f}  {e.name[:110]}\n")


def kernel_roofline(kb, stepper, math_mode, B, E):
    """Times the three tensor-tile kernels of a step in isolation (CUDA events on the launching stream, L2 flushed
    between launches; an empty label CSR so that only the tile kernel and its tiny pre/post kernels run) and reports the
    dominant one against the measured dense-bf16 peak.  `traffic` = dram bytes of that kernel from the committed
    ncu --set full capture (profiles/ncu_full_latest_summary.json)."""
    pk = peaks()
    st = stepper
    d = st.d
    dev = st.ent.device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lse = torch.zeros(B, device=dev)
    mp = None if st.mirror is None else st.mirror.data_ptr()
    L = kb.lib
    ent = st.ent.detach()
    gtmp = torch.zeros_like(ent)
    off0 = torch.zeros(B + 1, dtype=torch.int64, device=dev)   # no labels: isolates the dense tile kernels

    def fwd():
        L.call("kgeb_fused_fwd", st.loss_kind, math_mode, st.Q.data_ptr(), B, d, ent.data_ptr(), 0, E, E,
               off0.data_ptr(), st.lab_col.data_ptr(), 0, st.ls, st.offset, mp, st.rowstat.data_ptr(),
               st.ws.data_ptr(), st.ws.numel(), L.stream_ptr(ent))

    def bwd(dq, dt):
        L.call("kgeb_fused_bwd", st.loss_kind, math_mode, st.Q.data_ptr(), B, d, ent.data_ptr(), 0, E, E,
               off0.data_ptr(), st.lab_col.data_ptr(), 0, None, st.ls, st.offset, lse.data_ptr(), 1.0 / B, None,
               mp, st.dQ.data_ptr() if dq else None, gtmp.data_ptr() if dt else None, None, 0, st.ws.data_ptr(), st.ws.numel(),
               L.stream_ptr(ent))

    cases = {"tc_tiles_kernel<stats> (fused_fwd)": (fwd, 1, "tc::tc_tiles_kernel<1, 2, 1>"),
             "tc_bwd_kernel<dQ> (fused_bwd)": (lambda: bwd(True, False), 2, "tc_bwd_kernel<1, 1, 1, 0, 1>"),
             "tc_bwd_kernel<dTable> (fused_bwd)": (lambda: bwd(False, True), 2, "tc_bwd_kernel<0, 1, 1, 0, 0>")}
    res = {}
    for name, (fn, gemms, _) in cases.items():
        ts = []
        for i in range(8):
            flush.fill_(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        res[name] = (float(np.mean(ts)), gemms * 2.0 * B * E * d)
    name = max(res, key=lambda k: res[k][0])
    ms, flops = res[name]
    achieved = flops / (ms * 1e-3) / 1e12
    traffic = None
    path = os.path.join(ROOT, "profiles", "ncu_full_latest_summary.json")
    if os.path.exists(path):
        k = json.load(open(path)).get(cases[name][2])
        if k:
            traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
    tensor = math_mode != kb.lib.MATH_FP32
    return {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_burst"], "traffic": traffic,
            "note": ("algorithmic FLOPs = %d GEMM(s) x 2*B*E*d per launch (backward kernels recompute the score tile); "
                     "peak = %s dense bf16 burst; kernel timed alone incl. its bf16(Q) / reduce helpers, L2 flushed; "
                     "these fused kernels are MUFU-bound (ex2+rcp per score), see DESIGN.md 4.1; "
                     % (cases[name][1], pk["source"])) + ("tcgen05 tiles" if tensor else "CUDA-core fp32 tiles"),
            "all_ms": {k: v[0] for k, v in res.items()}}


if __name__ == "__main__":
    main()
