"""Peer-memory data-parallel step (csrc/p2p.cu) against the NCCL data-parallel step, run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/p2p_check.py

Both modes sum the ranks' gradients and apply one Adagrad step to replicated tables; the peer-memory mode does it in
one kernel per table (reduce-scatter of gradients in rank order, update of the owned slice, all-gather of the updated
weights) inside the step's CUDA graph.  At world 2 the sum of two terms is order-independent, so the two modes must
agree bit for bit; at larger worlds NCCL's reduction order differs and the comparison is to rounding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kgeb200 as kb  # noqa: E402
from oracle import kge_oracle as ko  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    g = kb.graph.synthetic_graph("toy", seed=3)
    e, r, d, b = g["num_entities"], g["num_relations"], 64, 96
    idx = [ko.kvsall_index(g["train"], "sp"), ko.kvsall_index(g["train"], "po")]
    rng = np.random.default_rng(100 + rank)
    batches = []
    for _ in range(4):
        ids = rng.choice(len(idx[0][0]) + len(idx[1][0]), b, replace=False)
        q, c, qt = ko.kvsall_collate(ids.tolist(), idx)
        batches.append({"queries": torch.from_numpy(q), "label_coords": torch.from_numpy(c),
                        "query_type_indexes": torch.from_numpy(qt)})
    nnz_max = max(len(x["label_coords"]) for x in batches)
    t = torch.tensor([nnz_max], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nnz_max = int(t.item())
    for math_mode in (kb.lib.MATH_BF16, kb.lib.MATH_FP32):
        torch.manual_seed(0)
        models = [kb.KgeModel("complex", e, r, d).to(dev) for _ in range(2)]
        models[1].load_state_dict(models[0].state_dict())
        jobs = []
        for m, p2p in zip(models, (False, True)):
            job = kb.TrainingJobKvsAll(m, kb.optim.create("Adagrad", m.parameters(), lr=0.2), kb.KgeLoss.create("bce"), e, r,
                                       math_mode=math_mode)
            job.enable_graph_step(b, nnz_max, use_graph=True, dp_group=dist.group.WORLD, dp_p2p=p2p)
            jobs.append(job)
        for i, batch in enumerate(batches):
            a, c = jobs[0].step(i, batch), jobs[1].step(i, batch)
            jobs[1].stepper.check_p2p()
            # world > 2: NCCL's ring / tree order differs from the rank order, and the difference compounds over the steps
            # (8 GPUs, step 3: 1.9e-6 relative)
            tol = 0.0 if world == 2 else 2e-5 * abs(a.avg_loss)
            assert abs(a.avg_loss - c.avg_loss) <= tol, (math_mode, i, a.avg_loss, c.avg_loss)
            for x, y in ((models[0].get_s_embedder().weight, models[1].get_s_embedder().weight),
                         (models[0].get_p_embedder().weight, models[1].get_p_embedder().weight)):
                err = (x - y).abs().max().item()
                assert err == 0.0 if world == 2 else err <= 2e-3, (math_mode, i, err)
        # replicas identical on every rank
        w = models[1].get_s_embedder().weight.detach()
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w.contiguous())
        assert all(torch.equal(ws[0], x) for x in ws), "replicas diverged"
        st = jobs[1].stepper
        if st.mirror is not None:
            assert torch.equal(st.mirror, w.bfloat16()), "bf16 mirror out of step with the table"
        st.sync_optimizer_state()
        s0 = jobs[0].optimizer.state[models[0].get_s_embedder().weight]["sum"]
        s1 = jobs[1].optimizer.state[models[1].get_s_embedder().weight]["sum"]
        err = (s0 - s1).abs().max().item()
        # (sum of g^2 over three steps; the two modes add the ranks' gradients in different orders: world - 1 roundings of 2^-24
        # per sum, squared and accumulated -- 1.4e-6 relative measured at world 8)
        assert err == 0.0 if world == 2 else err <= 5e-6 * s0.abs().max().item(), err
        if rank == 0:
            print(f"peer-memory data-parallel step == NCCL data-parallel step (world {world}, math={math_mode}); "
                  f"loss {c.avg_loss}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
