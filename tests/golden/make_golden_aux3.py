"""Golden vectors for the evaluation drill-down indexes and the argument-frequency metrics (build container only; writes
tests/golden/aux3.npz): `relation_types`, `relations_per_type`, `frequency_percentiles` (kge/indexing.py:142-263) on a
graph with all four relation types, and an EntityRankingJob run with entity_ranking.metrics_per.argument_frequency and
.relation_type (eval.py:173-224)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

E, R = mg.E, mg.R


def typed_graph(seed):
    """Relations 0-1 one-to-one, 2-3 one-to-many, 4 many-to-one, 5-6 many-to-many."""
    rng = np.random.default_rng(seed)
    t = []
    for p in (0, 1):
        perm = rng.permutation(E)
        t += [(s, p, perm[s]) for s in range(0, E, 2)]
    for p in (2, 3):
        for s in rng.choice(E, 12, replace=False):
            t += [(s, p, o) for o in rng.choice(E, 4, replace=False)]
    for o in rng.choice(E, 12, replace=False):
        t += [(s, 4, o) for s in rng.choice(E, 4, replace=False)]
    for p in (5, 6):
        s, o = rng.integers(0, 12, 120), rng.integers(0, 12, 120)
        t += list(zip(s.tolist(), [p] * 120, o.tolist()))
    t = np.unique(np.array(t, dtype=np.int32), axis=0)
    t = t[rng.permutation(len(t))]
    n = len(t)
    return {"train": t[: n - 80], "valid": t[n - 80: n - 40], "test": t[n - 40:]}


def main():
    out = {}
    g = typed_graph(11)
    for k, v in g.items():
        out[f"graph.{k}"] = v
    opts = {"train.type": "1vsAll", "eval.batch_size": 16, "entity_ranking.filter_with_test": True, "eval.split": "valid",
            "entity_ranking.metrics_per.head_and_tail": False, "entity_ranking.metrics_per.relation_type": True,
            "entity_ranking.metrics_per.argument_frequency": True, "entity_ranking.hits_at_k_s": [1, 3, 10]}
    job = mg.make_job("distmult", g, opts)
    with torch.no_grad():
        for emb in (job.model.get_s_embedder(), job.model.get_p_embedder()):
            w = emb._embeddings.weight
            w.copy_(torch.clamp(torch.round(w * 16 * 2.5) / 16, -0.5, 0.5))
    ds = job.dataset
    out["relation_types"] = np.array(ds.index("relation_types"))
    for t, rels in ds.index("relations_per_type").items():
        out[f"relations_of.{t}"] = np.array(sorted(int(r) for r in rels), dtype=np.int64)
    ds.index("frequency_percentiles")
    for arg, percs in ds._indexes["frequency_percentiles"].items():
        for perc, ids in percs.items():
            out[f"percentile.{arg}.{perc}"] = np.array(sorted(int(i) for i in ids), dtype=np.int64)
    ev = job.valid_job
    ev._prepare()
    trace = ev._run()
    out["ent"] = mg.np_(job.model.get_s_embedder()._embeddings.weight)
    out["rel"] = mg.np_(job.model.get_p_embedder()._embeddings.weight)
    keys = []
    for k, v in trace.items():
        if isinstance(v, float) and ("rank" in k or "hits" in k):
            out[f"value.{k}"] = np.float64(v)
            keys.append(k)
    out["keys"] = np.array(keys)
    path = os.path.join(HERE, "aux3.npz")
    np.savez_compressed(path, **out)
    print("aux3", len(out), "arrays", os.path.getsize(path), "bytes", "types", out["relation_types"])
    print(sorted(k for k in keys if "subject" in k or "1-N" in k)[:12])


if __name__ == "__main__":
    main()
