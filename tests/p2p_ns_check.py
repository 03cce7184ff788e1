"""Data-parallel negative-sampling step (FusedNegSamplingStepper(dp_group=...), trainer.PeerExchange, csrc/p2p.cu) against
the single-GPU step on the concatenated batch; run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/p2p_ns_check.py

NOT yet run on hardware (written after the GPU budget of round 1 was spent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kgeb200 as kb  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    g = kb.graph.synthetic_graph("toy", seed=3)
    e, r, d, b, n = g["num_entities"], g["num_relations"], 64, 32, 8
    rng = np.random.default_rng(7)                        # every rank builds all ranks' batches
    steps = 3
    triples = [[torch.from_numpy(g["train"][(s * world + k) * b:(s * world + k + 1) * b].astype(np.int64)) for k in range(world)]
               for s in range(steps)]
    negs = [[[torch.from_numpy(rng.integers(0, e, (b, n))), torch.zeros(b, 0, dtype=torch.long),
              torch.from_numpy(rng.integers(0, e, (b, n)))] for _ in range(world)] for _ in range(steps)]
    for model in ("transe", "complex"):
        torch.manual_seed(0)
        ref = kb.KgeModel(model, e, r, d).to(dev)
        new = kb.KgeModel(model, e, r, d).to(dev)
        new.load_state_dict(ref.state_dict())
        mk = lambda m: kb.optim.create("Adagrad", m.parameters(), lr=0.1, initial_accumulator_value=0.1)  # noqa: E731
        jr = kb.TrainingJobNegativeSampling(ref, mk(ref), kb.KgeLoss.create("kl"))
        jr.enable_graph_step(b * world, n, n, use_graph=False)
        jn = kb.TrainingJobNegativeSampling(new, mk(new), kb.KgeLoss.create("kl"))
        jn.enable_graph_step(b, n, n, use_graph=True, dp_group=dist.group.WORLD)
        for s in range(steps):
            glob = {"triples": torch.cat(triples[s]),
                    "negative_samples": [torch.cat([negs[s][k][slot] for k in range(world)]) for slot in range(3)]}
            mine = {"triples": triples[s][rank], "negative_samples": negs[s][rank]}
            a, c = jr.step(s, glob), jn.step(s, mine)
            jn.stepper.px.check()
            assert abs(a.avg_loss - c.avg_loss) <= 1e-5 * abs(a.avg_loss), (model, s, a.avg_loss, c.avg_loss)
            for x, y in ((ref.get_s_embedder().weight, new.get_s_embedder().weight),
                         (ref.get_p_embedder().weight, new.get_p_embedder().weight)):
                err = (x - y).abs().max().item()
                assert err <= 1e-5, (model, s, err)
        w = new.get_s_embedder().weight.detach().contiguous()
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        assert all(torch.equal(ws[0], x) for x in ws), "replicas diverged"
        if rank == 0:
            print(f"data-parallel negative-sampling step == single-GPU step on the global batch ({model}, world {world}); "
                  f"loss {c.avg_loss}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
