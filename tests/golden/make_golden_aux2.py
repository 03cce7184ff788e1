"""More golden vectors from the UNMODIFIED reference (build container only; writes tests/golden/aux2.npz):

  * ReciprocalRelationsModel over distmult / complex (reciprocal_relations_model.py:56-106): score_spo for both
    directions, score_sp, score_po, score_sp_po (+ entity subset), and one 1vsAll training step;
  * two training steps with torch.optim.Adam (KgeOptimizer.create, util/optimizer.py:10-17);
  * the toy example as written (examples/toy-complex-train.yaml: ComplEx, KvsAll + KL, regularize_weight 0.8e-7 -- here
    1e-2 so that the penalty is visible): two steps of run_epoch's body with the penalty terms back-propagated between the
    loss and optimizer.step() (train.py:309-376).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the import shims)

E, R, D = mg.E, mg.R, mg.D


def make_reciprocal_job(base, graph, options):
    """mg.make_job for reciprocal_relations_model: the base model's defaults must be imported as well."""
    import tempfile
    from kge import Config, Dataset
    from kge.job import Job
    tmp = tempfile.mkdtemp()
    c = Config(folder=tmp)
    c.set("model", "reciprocal_relations_model")
    c._import("reciprocal_relations_model")
    c._import(base)
    c.set("reciprocal_relations_model.base_model.type", base)
    c.set("dataset.num_entities", E)
    c.set("dataset.num_relations", R)
    c.set("dataset.pickle", False)
    c.set("job.device", "cpu")
    c.set("lookup_embedder.dim", D)
    c.set("train.optimizer", "Adagrad")
    c.set("train.optimizer_args", {"lr": 0.2}, create=True)
    for k, v in options.items():
        c.set(k, v, create=True)
    ds = Dataset(c, folder=tmp)
    for split, arr in graph.items():
        ds._triples[split] = torch.from_numpy(arr.copy())
    torch.manual_seed(0)
    job = Job.create(c, ds)
    c.log = lambda *a, **k: None
    return job


def gen_reciprocal(out):
    g = mg.make_graph(5)
    rng = np.random.default_rng(6)
    n = 9
    s, p, o = rng.integers(0, E, n), rng.integers(0, R, n), rng.integers(0, E, n)
    subset = rng.permutation(E)[:17]
    out["recip.idx_s"], out["recip.idx_p"], out["recip.idx_o"], out["recip.subset"] = s, p, o, subset
    out["recip.graph.train"] = g["train"]
    names = []
    for base in ("distmult", "complex"):
        tag = f"recip.{base}"
        names.append(tag)
        opts = {"train.type": "1vsAll", "train.loss": "kl", "train.batch_size": 16}
        job = make_reciprocal_job(base, g, opts)
        mod = job.model
        st, pt, ot = (torch.from_numpy(x) for x in (s, p, o))
        with torch.no_grad():
            out[tag + ".ent"] = mg.np_(mod.get_s_embedder()._embeddings.weight)
            out[tag + ".rel"] = mg.np_(mod.get_p_embedder()._embeddings.weight)      # [2R, d]
            out[tag + ".spo_o"] = mg.np_(mod.score_spo(st, pt, ot, "o"))
            out[tag + ".spo_s"] = mg.np_(mod.score_spo(st, pt, ot, "s"))
            out[tag + ".sp"] = mg.np_(mod.score_sp(st, pt))
            out[tag + ".po"] = mg.np_(mod.score_po(pt, ot))
            out[tag + ".sp_po"] = mg.np_(mod.score_sp_po(st, pt, ot))
            out[tag + ".sp_po_sub"] = mg.np_(mod.score_sp_po(st, pt, ot, torch.from_numpy(subset)))
        job._prepare()
        torch.manual_seed(11)
        batch = next(iter(job.loader))
        out[tag + ".b0.triples"] = mg.np_(batch["triples"])
        lv, ge, gr = mg.run_train_step(job, batch)
        out[tag + ".b0.loss"] = np.float64(lv)
        out[tag + ".b0.grad_ent"], out[tag + ".b0.grad_rel"] = ge, gr
        out[tag + ".b0.ent"] = mg.np_(mod.get_s_embedder()._embeddings.weight)
        out[tag + ".b0.rel"] = mg.np_(mod.get_p_embedder()._embeddings.weight)
    out["recip.cases"] = np.array(names)


def gen_adam(out):
    g = mg.make_graph(3)
    opts = {"train.type": "1vsAll", "train.loss": "kl", "train.batch_size": 16, "train.optimizer": "Adam",
            "train.optimizer_args": {"lr": 0.01}}
    job = mg.make_job("distmult", g, opts)
    job._prepare()
    out["adam.graph.train"] = g["train"]
    out["adam.ent0"] = mg.np_(job.model.get_s_embedder()._embeddings.weight)
    out["adam.rel0"] = mg.np_(job.model.get_p_embedder()._embeddings.weight)
    torch.manual_seed(11)
    it = iter(job.loader)
    for step in range(2):
        batch = next(it)
        out[f"adam.b{step}.triples"] = mg.np_(batch["triples"])
        lv, ge, gr = mg.run_train_step(job, batch)
        out[f"adam.b{step}.loss"] = np.float64(lv)
        out[f"adam.b{step}.grad_ent"], out[f"adam.b{step}.grad_rel"] = ge, gr
        out[f"adam.b{step}.ent"] = mg.np_(job.model.get_s_embedder()._embeddings.weight)
        out[f"adam.b{step}.rel"] = mg.np_(job.model.get_p_embedder()._embeddings.weight)
    st = job.optimizer.state[job.model.get_s_embedder()._embeddings.weight]
    out["adam.exp_avg_ent"], out["adam.exp_avg_sq_ent"] = mg.np_(st["exp_avg"]), mg.np_(st["exp_avg_sq"])


def gen_toy_penalty(out):
    g = mg.make_graph(3)
    opts = {"train.type": "KvsAll", "train.loss": "kl", "train.batch_size": 16,
            "lookup_embedder.regularize_weight": 1e-2}
    job = mg.make_job("complex", g, opts)
    job._prepare()
    model = job.model
    out["toy.graph.train"] = g["train"]
    out["toy.ent0"] = mg.np_(model.get_s_embedder()._embeddings.weight)
    out["toy.rel0"] = mg.np_(model.get_p_embedder()._embeddings.weight)
    torch.manual_seed(11)
    it = iter(job.loader)
    for step in range(2):
        batch = next(it)
        out[f"toy.b{step}.queries"] = mg.np_(batch["queries"])
        out[f"toy.b{step}.label_coords"] = mg.np_(batch["label_coords"])
        out[f"toy.b{step}.query_type"] = mg.np_(batch["query_type_indexes"])
        # run_epoch's body, train.py:309-376
        job.optimizer.zero_grad()
        res = job._process_batch(step, batch)
        terms = model.penalty(epoch=1, batch_index=step, num_batches=2, batch=batch)
        vals = []
        for _, v in terms:
            v.backward()
            vals.append(v.item())
        out[f"toy.b{step}.loss"] = np.float64(res.avg_loss)
        out[f"toy.b{step}.penalties"] = np.array(vals, dtype=np.float64)
        out[f"toy.b{step}.grad_ent"] = mg.np_(model.get_s_embedder()._embeddings.weight.grad)
        out[f"toy.b{step}.grad_rel"] = mg.np_(model.get_p_embedder()._embeddings.weight.grad)
        job.optimizer.step()
        out[f"toy.b{step}.ent"] = mg.np_(model.get_s_embedder()._embeddings.weight)
        out[f"toy.b{step}.rel"] = mg.np_(model.get_p_embedder()._embeddings.weight)
    out["toy.penalty_keys"] = np.array([k for k, _ in terms])


def main():
    out = {}
    gen_reciprocal(out)
    gen_adam(out)
    gen_toy_penalty(out)
    path = os.path.join(HERE, "aux2.npz")
    np.savez_compressed(path, **out)
    print("aux2", len(out), "arrays", os.path.getsize(path), "bytes")
    print({k: out[k] for k in out if k.endswith(".loss") or k.endswith("penalties")})


if __name__ == "__main__":
    main()
