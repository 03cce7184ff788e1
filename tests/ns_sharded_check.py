"""Negative sampling data-parallel over the triples with the entity table SHARDED BY ROW (second row:
trainer.ShardedNegSamplingStepper -- all-to-all-v of requested rows forward, of row gradients backward, owner-side
deterministic segment sum + Adagrad) against the single-GPU step on the concatenated batch; run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tests/ns_sharded_check.py
"""
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
    for k in range(world):       # a hub entity and a row every rank asks for: long segments at one owner
        negs[0][k][0][:, :3] = 5
        negs[0][k][2][:, 0] = e - 1
    for model, loss in (("rotate", "kl"), ("transe", "bce"), ("complex", "kl"), ("distmult", "bce")):
        torch.manual_seed(0)
        ref = kb.KgeModel(model, e, r, d).to(dev)
        new = kb.KgeModel(model, e, r, d).to(dev)
        new.load_state_dict(ref.state_dict())
        mk = lambda m: kb.optim.create("Adagrad", m.parameters(), lr=0.1, initial_accumulator_value=0.1)  # noqa: E731
        jr = kb.TrainingJobNegativeSampling(ref, mk(ref), kb.KgeLoss.create(loss))
        jr.enable_graph_step(b * world, n, n, use_graph=False)
        kind = kb.lib.LOSS_KL if loss == "kl" else kb.lib.LOSS_BCE
        st = kb.trainer.ShardedNegSamplingStepper(new, mk(new), b, n, n, kind, dist.group.WORLD)
        for s in range(steps):
            glob = {"triples": torch.cat(triples[s]),
                    "negative_samples": [torch.cat([negs[s][k][slot] for k in range(world)]) for slot in range(3)]}
            a = jr.step(s, glob)
            c = st.step(triples[s][rank], negs[s][rank]).item()
            assert abs(a.avg_loss - c) <= 1e-5 * abs(a.avg_loss), (model, s, a.avg_loss, c)
        st.sync_tables()
        for x, y, what in ((ref.get_s_embedder().weight, new.get_s_embedder().weight, "entity table"),
                           (ref.get_p_embedder().weight, new.get_p_embedder().weight, "relation table"),
                           (jr.optimizer.state[ref.get_s_embedder().weight]["sum"], st.opt.state[new.get_s_embedder().weight]["sum"], "Adagrad state")):
            err = (x - y).abs().max().item()
            assert err <= 2e-5, (model, what, err)
        w = new.get_s_embedder().weight.detach().contiguous()
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        assert all(torch.equal(ws[0], x) for x in ws), "tables differ after sync_tables()"
        if rank == 0:
            print(f"row-sharded data-parallel negative-sampling step == single-GPU step on the global batch ({model}, {loss}, "
                  f"world {world}); loss {c}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
