"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

imports /root/reference through the shims in ref_import.py, drives the reference's own
KgeModel / TrainingJob* / EntityRankingJob / KvsAllIndex code on small seeded synthetic graphs
(CPU, fp32) and writes inputs + outputs to tests/golden/*.npz.  The oracle (oracle/kge_oracle.py)
and, on the GPU box, the CUDA path are compared against these files; the reference tree itself
does not travel to the GPU box.
"""
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

warnings.filterwarnings("ignore")
ref_import.install()

from kge import Config, Dataset  # noqa: E402
from kge.job import Job  # noqa: E402
from kge.job.util import get_sp_po_coords_from_spo_batch  # noqa: E402
from kge.job.entity_ranking import EntityRankingJob  # noqa: E402

E, R, D = 53, 7, 16
MODELS = ["distmult", "complex", "cp", "simple", "rescal", "transe", "rotate"]


def make_graph(seed, e=E, r=R, n_train=400, n_valid=40, n_test=40):
    rng = np.random.default_rng(seed)
    n = n_train + n_valid + n_test
    # skewed draws so that (s,p) / (p,o) keys have several answers
    s = np.minimum(rng.zipf(1.3, n) - 1, e - 1)
    o = np.minimum(rng.zipf(1.3, n) - 1, e - 1)
    perm = rng.permutation(e)
    p = rng.integers(0, r, n)
    t = np.unique(np.stack([perm[s], p, perm[(o + 7) % e]], 1), axis=0)
    t = t[rng.permutation(len(t))].astype(np.int32)
    n_train = len(t) - n_valid - n_test
    return {"train": t[:n_train], "valid": t[n_train:n_train + n_valid], "test": t[n_train + n_valid:]}


def make_job(model, graph, options, e=E, r=R, d=D):
    tmp = tempfile.mkdtemp()
    c = Config(folder=tmp)
    c.set("model", model)
    c._import(model)
    c.set("dataset.num_entities", e)
    c.set("dataset.num_relations", r)
    c.set("dataset.pickle", False)
    c.set("job.device", "cpu")
    c.set("lookup_embedder.dim", d)
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


def np_(t):
    return t.detach().cpu().numpy().copy()


def gen_scores(out):
    g = make_graph(1)
    rng = np.random.default_rng(5)
    n = 11
    s = rng.integers(0, E, n)
    p = rng.integers(0, R, n)
    o = rng.integers(0, E, n)
    subset = rng.permutation(E)[:19]
    out["idx_s"], out["idx_p"], out["idx_o"], out["subset"] = s, p, o, subset
    for m in MODELS:
        for ln in ([1.0, 2.0] if m in ("transe", "rotate") else [None]):
            opts = {"train.type": "1vsAll"}
            tag = m
            if ln is not None:
                opts[m + ".l_norm"] = ln
                tag = f"{m}_l{int(ln)}"
            job = make_job(m, g, opts)
            mod = job.model
            st, pt, ot = (torch.from_numpy(x) for x in (s, p, o))
            with torch.no_grad():
                out[f"{tag}.ent"] = np_(mod.get_s_embedder()._embeddings.weight)
                out[f"{tag}.rel"] = np_(mod.get_p_embedder()._embeddings.weight)
                out[f"{tag}.spo"] = np_(mod.score_spo(st, pt, ot))
                out[f"{tag}.sp"] = np_(mod.score_sp(st, pt))
                out[f"{tag}.po"] = np_(mod.score_po(pt, ot))
                out[f"{tag}.sp_sub"] = np_(mod.score_sp(st, pt, torch.from_numpy(subset)))
                out[f"{tag}.sp_po_sub"] = np_(mod.score_sp_po(st, pt, ot, torch.from_numpy(subset)))
                out[f"{tag}.so"] = np_(mod.score_so(st, ot))
                # int32 indexes as in evaluation (dataset.py:178)
                out[f"{tag}.spo_i32"] = np_(mod.score_spo(st.int(), pt.int(), ot.int()))


def gen_losses(out):
    from kge.util.loss import KgeLoss
    rng = np.random.default_rng(9)
    sc = torch.from_numpy(rng.normal(0, 2, (9, 23)).astype(np.float32))
    idx = torch.from_numpy(rng.integers(0, 23, 9))
    mat = torch.from_numpy((rng.random((9, 23)) < 0.15).astype(np.float32))
    mat[torch.arange(9), idx] = 1.0
    out["loss.scores"], out["loss.idx"], out["loss.mat"] = np_(sc), np_(idx), np_(mat)
    for name, arg in [("kl", float("nan")), ("bce", float("nan")), ("bce", 0.7)]:
        c = Config()
        c.log = lambda *a, **k: None
        c.set("train.loss", name)
        c.set("train.loss_arg", arg)
        c.set("job.device", "cpu")
        loss = KgeLoss.create(c)
        tag = f"loss.{name}" + ("" if arg != arg else "_off")
        out[tag + ".idx"] = np.float64(loss(sc, idx).item())
        out[tag + ".mat"] = np.float64(loss(sc, mat).item())


def gen_index(out):
    g = make_graph(2)
    job = make_job("complex", g, {"train.type": "KvsAll"})
    ds = job.dataset
    for split in ("train", "valid"):
        out[f"index.{split}.triples"] = g[split]
        for key, val in (("sp", "o"), ("po", "s")):
            ix = ds.index(f"{split}_{key}_to_{val}")
            out[f"index.{split}.{key}.keys"] = np_(ix._keys)
            out[f"index.{split}.{key}.offsets"] = np_(ix._values_offset)
            out[f"index.{split}.{key}.values"] = np_(ix._values)
    batch = [torch.from_numpy(x) for x in g["valid"][:13]]
    coords = get_sp_po_coords_from_spo_batch(batch, E, ds.index("train_sp_to_o"), ds.index("train_po_to_s"))
    out["index.coords_batch"] = g["valid"][:13]
    out["index.coords"] = np_(coords)


def run_train_step(job, batch):
    job.optimizer.zero_grad()
    res = job._process_batch(0, batch)
    ge = np_(job.model.get_s_embedder()._embeddings.weight.grad)
    gr = np_(job.model.get_p_embedder()._embeddings.weight.grad)
    job.optimizer.step()
    return res.avg_loss, ge, gr


def gen_train(out):
    g = make_graph(3)
    out["train.graph.train"] = g["train"]
    cases = [
        ("1vsAll", "distmult", "kl", {}),
        ("1vsAll", "complex", "kl", {}),
        ("1vsAll", "rescal", "bce", {}),
        ("1vsAll", "transe", "kl", {}),
        ("KvsAll", "complex", "bce", {}),
        ("KvsAll", "complex", "kl", {}),
        ("KvsAll", "simple", "bce", {"KvsAll.label_smoothing": 0.1, "train.loss_arg": 0.5}),
        ("KvsAll", "cp", "kl", {}),
        ("KvsAll", "rotate", "kl", {}),  # rotate _po + bce fails in the reference (.t() then .view, loss.py:157)
        ("negative_sampling", "rotate", "kl", {"negative_sampling.num_samples.s": 5,
                                               "negative_sampling.num_samples.o": 6}),
        ("negative_sampling", "transe", "bce", {"negative_sampling.num_samples.s": 4, "transe.l_norm": 2.0}),
        ("negative_sampling", "distmult", "kl", {"negative_sampling.num_samples.s": 3}),
        ("negative_sampling", "complex", "bce", {"negative_sampling.num_samples.s": 3,
                                                 "negative_sampling.num_samples.o": 2}),
    ]
    names = []
    for ttype, model, loss, extra in cases:
        tag = f"train.{ttype}.{model}.{loss}"
        names.append(tag)
        opts = {"train.type": ttype, "train.loss": loss, "train.batch_size": 16,
                "negative_sampling.implementation": "triple"}
        opts.update(extra)
        job = make_job(model, g, opts)
        job._prepare()
        out[tag + ".ent0"] = np_(job.model.get_s_embedder()._embeddings.weight)
        out[tag + ".rel0"] = np_(job.model.get_p_embedder()._embeddings.weight)
        torch.manual_seed(11)
        it = iter(job.loader)
        for step in range(2):  # two optimizer steps: exercises the Adagrad accumulator
            batch = next(it)
            if ttype == "KvsAll":
                out[f"{tag}.b{step}.queries"] = np_(batch["queries"])
                out[f"{tag}.b{step}.label_coords"] = np_(batch["label_coords"])
                out[f"{tag}.b{step}.query_type"] = np_(batch["query_type_indexes"])
            else:
                out[f"{tag}.b{step}.triples"] = np_(batch["triples"])
                if ttype == "negative_sampling":
                    for slot, ns in enumerate(batch["negative_samples"]):
                        out[f"{tag}.b{step}.neg{slot}"] = np_(ns)
            lv, ge, gr = run_train_step(job, batch)
            out[f"{tag}.b{step}.loss"] = np.float64(lv)
            out[f"{tag}.b{step}.grad_ent"] = ge
            out[f"{tag}.b{step}.grad_rel"] = gr
            out[f"{tag}.b{step}.ent"] = np_(job.model.get_s_embedder()._embeddings.weight)
            out[f"{tag}.b{step}.rel"] = np_(job.model.get_p_embedder()._embeddings.weight)
        if extra:
            out[tag + ".options"] = np.array(repr(sorted(extra.items())))
    out["train.cases"] = np.array(names)


def gen_eval(out):
    g = make_graph(4, n_train=500, n_valid=48, n_test=48)
    for k, v in g.items():
        out[f"eval.graph.{k}"] = v
    names = []
    for model, chunk, ties, extra in [
        ("complex", -1, "rounded_mean_rank", {}),
        ("complex", 17, "worst_rank", {}),
        ("transe", 20, "rounded_mean_rank", {}),
        ("rotate", -1, "best_rank", {}),
        ("distmult", 11, "rounded_mean_rank", {}),
    ]:
        tag = f"eval.{model}.c{chunk}.{ties}"
        names.append(tag)
        opts = {"train.type": "1vsAll", "eval.batch_size": 16, "entity_ranking.chunk_size": chunk,
                "entity_ranking.tie_handling": ties, "entity_ranking.filter_with_test": True,
                "eval.split": "valid"}
        opts.update(extra)
        job = make_job(model, g, opts)
        with torch.no_grad():
            # coarse dyadic weights (multiples of 1/16, |w| <= 1/2) => plenty of exact score ties, and every
            # product / sum is exact in fp32 (and in TF32) whatever the summation order, so tie counts are
            # reproducible bit-for-bit by any implementation (rotate: sin/cos/sqrt make ties a matter of chance)
            for emb in (job.model.get_s_embedder(), job.model.get_p_embedder()):
                w = emb._embeddings.weight
                w.copy_(torch.clamp(torch.round(w * 16 * 2.5) / 16, -0.5, 0.5) if model != "rotate"
                        else torch.round(w * 4) / 4)
        ev = job.valid_job
        ev._prepare()
        rec = []

        def hook(hists, s, p, o, s_ranks, o_ranks, job=None, **kw):
            rec.append((np_(s_ranks), np_(o_ranks)))

        ev.hist_hooks.append(hook)
        trace = ev._run()
        out[tag + ".ent"] = np_(job.model.get_s_embedder()._embeddings.weight)
        out[tag + ".rel"] = np_(job.model.get_p_embedder()._embeddings.weight)
        # hooks are called raw, filt, filt_test per batch (entity_ranking.py:234-268)
        for j, nm in enumerate(["raw", "filt", "filt_test"]):
            out[f"{tag}.s_ranks_{nm}"] = np.concatenate([rec[i][0] for i in range(j, len(rec), 3)])
            out[f"{tag}.o_ranks_{nm}"] = np.concatenate([rec[i][1] for i in range(j, len(rec), 3)])
        for k, v in trace.items():
            if isinstance(v, float) and ("rank" in k or "hits" in k):
                out[f"{tag}.metric.{k}"] = np.float64(v)
    out["eval.cases"] = np.array(names)
    # the reference's only known-answer vector: config-default.yaml:392-406
    sc = torch.tensor([[10.0, 10.0, 10.0, 11.0, 9.0]])
    rank, ties = EntityRankingJob._get_ranks_and_num_ties(sc, torch.tensor([10.0]))
    out["eval.tie_example.rank_ties"] = np.array([rank.item(), ties.item()])


def main():
    for name, fn in [("scores", gen_scores), ("losses", gen_losses), ("index", gen_index),
                     ("train", gen_train), ("eval", gen_eval)]:
        out = {}
        fn(out)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
