"""Golden vectors for the rows around the hot path, from the UNMODIFIED reference:

    python tests/golden/make_golden_aux.py      (build container only; writes tests/golden/aux.npz)

  * Lp penalties: KgeModel.penalty (kge_model.py:588-606) -> LookupEmbedder.penalty (lookup_embedder.py:112-158),
    values and the parameter gradients after back-propagating every term (train.py:320-338);
  * ranking metrics with the head/tail and per-relation-type drill-downs (eval.py:138-198,
    entity_ranking.py:370-381), plus the reference's relations_per_type index.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the import shims)


def gen_penalty(out):
    g = mg.make_graph(3)
    triples = torch.from_numpy(g["train"][:24].astype(np.int64))
    out["penalty.triples"] = mg.np_(triples)
    names = []
    for p in (1, 2, 3):
        for weighted in (False, True):
            tag = f"penalty.p{p}.{'weighted' if weighted else 'unweighted'}"
            names.append(tag)
            opts = {"train.type": "negative_sampling"}
            for emb in ("entity_embedder", "relation_embedder"):
                opts[f"complex.{emb}.regularize_weight"] = 0.05 if emb == "entity_embedder" else 0.02
                opts[f"complex.{emb}.regularize_args.p"] = p
                opts[f"complex.{emb}.regularize_args.weighted"] = weighted
            job = mg.make_job("complex", g, opts)
            model = job.model
            ent, rel = model.get_s_embedder()._embeddings.weight, model.get_p_embedder()._embeddings.weight
            out[tag + ".ent"], out[tag + ".rel"] = mg.np_(ent), mg.np_(rel)
            terms = model.penalty(epoch=1, batch_index=0, num_batches=1, batch={"triples": triples})
            vals = []
            for _, v in terms:
                v.backward()
                vals.append(v.item())
            out[tag + ".values"] = np.array(vals, dtype=np.float64)
            out[tag + ".keys"] = np.array([k for k, _ in terms])
            out[tag + ".grad_ent"], out[tag + ".grad_rel"] = mg.np_(ent.grad), mg.np_(rel.grad)
    out["penalty.cases"] = np.array(names)


def gen_metrics(out):
    g = mg.make_graph(4, n_train=500, n_valid=48, n_test=48)
    for k, v in g.items():
        out[f"metrics.graph.{k}"] = v
    opts = {"train.type": "1vsAll", "eval.batch_size": 16, "entity_ranking.filter_with_test": True,
            "eval.split": "valid", "entity_ranking.metrics_per.head_and_tail": True,
            "entity_ranking.metrics_per.relation_type": True, "entity_ranking.hits_at_k_s": [1, 3, 10]}
    job = mg.make_job("complex", g, opts)
    with torch.no_grad():
        for emb in (job.model.get_s_embedder(), job.model.get_p_embedder()):
            w = emb._embeddings.weight
            w.copy_(torch.clamp(torch.round(w * 16 * 2.5) / 16, -0.5, 0.5))
    ev = job.valid_job
    ev._prepare()
    trace = ev._run()
    out["metrics.ent"] = mg.np_(job.model.get_s_embedder()._embeddings.weight)
    out["metrics.rel"] = mg.np_(job.model.get_p_embedder()._embeddings.weight)
    rpt = job.dataset.index("relations_per_type")
    out["metrics.relation_types"] = np.array(sorted(rpt.keys()))
    for t, rels in rpt.items():
        out[f"metrics.relations_of.{t}"] = np.array(sorted(int(r) for r in rels), dtype=np.int64)
    keys = []
    for k, v in trace.items():
        if isinstance(v, float) and ("rank" in k or "hits" in k):
            out[f"metrics.value.{k}"] = np.float64(v)
            keys.append(k)
    out["metrics.keys"] = np.array(keys)


def main():
    out = {}
    gen_penalty(out)
    gen_metrics(out)
    path = os.path.join(HERE, "aux.npz")
    np.savez_compressed(path, **out)
    print("aux", len(out), "arrays", os.path.getsize(path), "bytes")
    print(sorted(k for k in out if k.startswith("metrics.value."))[:40])
    print({k: out[k] for k in out if k.endswith(".values")})


if __name__ == "__main__":
    main()
