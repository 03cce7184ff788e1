"""Synthetic knowledge graphs of the shapes named in BASELINE.json: subjects and
objects from a truncated Zipf over a random permutation of entity ids, relations from a Zipf over
relation ids, exact duplicates removed, disjoint splits, int32 [N,3] (s,p,o) as in dataset.py:178."""
from __future__ import annotations

from typing import Dict

import numpy as np

SHAPES = {
    "toy": dict(num_entities=280, num_relations=112, train=4565, valid=109, test=152),
    "fb15k-237": dict(num_entities=14541, num_relations=237, train=272115, valid=17535, test=20466),
    "wnrr": dict(num_entities=40943, num_relations=11, train=86835, valid=3034, test=3134),
    "wikidata5m": dict(num_entities=4_600_000, num_relations=822, train=20_000_000, valid=5000, test=5000),
}


def _zipf_ids(rng, n, size, alpha=1.0):
    # truncated Zipf via inverse-CDF on ranks 1..n
    w = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** alpha
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(size), side="left").astype(np.int64)


def synthetic_graph(shape: str = "fb15k-237", seed: int = 0, scale: float = 1.0) -> Dict[str, object]:
    """Returns {"num_entities", "num_relations", "train", "valid", "test"}; `scale` shrinks the triple counts."""
    cfg = SHAPES[shape]
    e, r = cfg["num_entities"], cfg["num_relations"]
    want = {k: max(1, int(cfg[k] * scale)) for k in ("train", "valid", "test")}
    total = sum(want.values())
    rng = np.random.default_rng(seed)
    perm = rng.permutation(e)
    out = np.zeros((0, 3), dtype=np.int64)
    while len(out) < total:
        n = int((total - len(out)) * 1.3) + 16
        s = perm[_zipf_ids(rng, e, n)]
        o = perm[(_zipf_ids(rng, e, n) + rng.integers(1, e, n)) % e] if shape == "none" else perm[_zipf_ids(rng, e, n)]
        p = _zipf_ids(rng, r, n)
        cand = np.stack([s, p, o], 1)
        out = np.concatenate([out, cand])
        # remove exact duplicates, keep first occurrence order
        key = (out[:, 0] * r + out[:, 1]) * e + out[:, 2]
        _, first = np.unique(key, return_index=True)
        out = out[np.sort(first)]
    out = out[rng.permutation(len(out))[:total]].astype(np.int32)
    a, b = want["train"], want["train"] + want["valid"]
    return {"num_entities": e, "num_relations": r, "train": out[:a], "valid": out[a:b], "test": out[b:total]}
