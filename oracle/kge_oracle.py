"""CPU oracle for the LibKGE scoring/embedding hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU fp32 restatement of the reference's algorithm for the hot path
named by BASELINE.json (gather -> score spo/sp_/_po -> KL/BCE loss -> backward -> Adagrad row
update, and filtered entity ranking).  The reference (Nzteb/kge-1, a LibKGE fork) is pure
Python whose arithmetic lives in the third-party dependency PyTorch (`torch>=1.3.1`,
reference setup.py:10; this image has torch 2.11.0).  The oracle therefore calls the same ATen
CPU ops at the same call sites (mm / bmm / cdist / pairwise_distance / norm / embedding /
BCEWithLogits / CrossEntropy / KLDiv / torch.optim.Adagrad) in a function-style layout of its
own; every function cites the reference file:line it follows.

PINNING: the oracle is checked against golden vectors produced by the *unmodified reference*
imported in the build container (tests/golden/make_golden.py -> tests/golden/*.npz, see
tests/test_oracle_golden.py) and against the reference's only known-answer vector for this
path, the tie-handling example in kge/config-default.yaml:392-406.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (kge-1_b200/) never does.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

S, P, O = 0, 1, 2
MODELS = ("distmult", "complex", "cp", "simple", "rescal", "transe", "rotate")


# --------------------------------------------------------------------------------------
# dimensions
# --------------------------------------------------------------------------------------
def relation_dim(model: str, d: int) -> int:
    """Relation embedding width for entity width d.

    cp.py:50-55 and rotate.py:82-87 (d/2), rescal.py:79-96 (d*d), otherwise lookup_embedder dim.
    """
    if model in ("cp", "rotate"):
        return d // 2
    if model == "rescal":
        return d * d
    return d


# --------------------------------------------------------------------------------------
# a1/a2: LookupEmbedder.embed / embed_all  (embedder/lookup_embedder.py:91-107), dropout == 0
# --------------------------------------------------------------------------------------
def embed(weight: torch.Tensor, indexes: torch.Tensor) -> torch.Tensor:
    return F.embedding(indexes.long(), weight)


def embed_all(weight: torch.Tensor) -> torch.Tensor:
    return F.embedding(torch.arange(weight.shape[0], dtype=torch.long), weight)


# --------------------------------------------------------------------------------------
# a8-a14: RelationalScorer.score_emb for the seven scorers
# --------------------------------------------------------------------------------------
def _generic_score_emb(model, s_emb, p_emb, o_emb, combine, l_norm):
    """kge_model.py:152-182: repeat/interleave fallback on top of the row-wise "spo" form."""
    n = p_emb.size(0)
    if combine == "sp_":
        n_o = o_emb.size(0)
        out = score_emb(model, s_emb.repeat_interleave(n_o, 0), p_emb.repeat_interleave(n_o, 0),
                        o_emb.repeat((n, 1)), "spo", l_norm)
    elif combine == "_po":
        n_s = s_emb.size(0)
        out = score_emb(model, s_emb.repeat((n, 1)), p_emb.repeat_interleave(n_s, 0),
                        o_emb.repeat_interleave(n_s, 0), "spo", l_norm)
    elif combine == "s_o":
        n = s_emb.size(0)
        n_p = p_emb.size(0)
        out = score_emb(model, s_emb.repeat_interleave(n_p, 0), p_emb.repeat((n, 1)),
                        o_emb.repeat_interleave(n_p, 0), "spo", l_norm)
    else:
        raise ValueError('cannot handle combine="{}"'.format(combine))
    return out.view(n, -1)


def _rot_abs(re, im):
    # rotate.py:142-145  norm over a stacked (re, im) axis
    return torch.norm(torch.stack((re, im), dim=0), dim=0)


def score_emb(model: str, s_emb, p_emb, o_emb, combine: str, l_norm: float = 1.0):
    n = p_emb.size(0)
    if combine not in ("spo", "sp_", "_po"):
        out = _generic_score_emb(model, s_emb, p_emb, o_emb, combine, l_norm)
        if model in ("rescal", "transe"):
            # rescal.py:50-53 and transe.py:22-24 re-view the fallback result with
            # n = p_emb.size(0); for "s_o" this reshapes the [n, R] result to [R, n]
            # (a reference quirk the oracle keeps)
            out = out.view(n, -1)
        return out

    if model == "distmult":  # distmult.py:14-26
        if combine == "spo":
            out = (s_emb * p_emb * o_emb).sum(dim=1)
        elif combine == "sp_":
            out = (s_emb * p_emb).mm(o_emb.t())
        else:
            out = (o_emb * p_emb).mm(s_emb.t())
    elif model == "complex":  # complex.py:19-44 (4-block concatenation, contraction width 2d)
        p_re, p_im = (t.contiguous() for t in p_emb.chunk(2, dim=1))
        o_re, o_im = (t.contiguous() for t in o_emb.chunk(2, dim=1))
        s_all = torch.cat((s_emb, s_emb), dim=1)
        r_all = torch.cat((p_re, p_emb, -p_im), dim=1)
        o_all = torch.cat((o_emb, o_im, o_re), dim=1)
        if combine == "spo":
            out = (s_all * o_all * r_all).sum(dim=1)
        elif combine == "sp_":
            out = (s_all * r_all).mm(o_all.t())
        else:
            out = (r_all * o_all).mm(s_all.t())
    elif model == "cp":  # cp.py:14-31
        h = s_emb.shape[1] // 2
        s_h, o_t = s_emb[:, :h], o_emb[:, h:]
        if combine == "spo":
            out = (s_h * p_emb * o_t).sum(dim=1)
        elif combine == "sp_":
            out = (s_h * p_emb).mm(o_t.t())
        else:
            out = (o_t * p_emb).mm(s_h.t())
    elif model == "simple":  # simple.py:14-34
        s_h, s_t = torch.chunk(s_emb, 2, dim=1)
        p_f, p_b = torch.chunk(p_emb, 2, dim=1)
        o_h, o_t = torch.chunk(o_emb, 2, dim=1)
        if combine == "spo":
            a = (s_h * p_f * o_t).sum(dim=1)
            b = (s_t * p_b * o_h).sum(dim=1)
        elif combine == "sp_":
            a = (s_h * p_f).mm(o_t.t())
            b = (s_t * p_b).mm(o_h.t())
        else:
            a = (o_t * p_f).mm(s_h.t())
            b = (o_h * p_b).mm(s_t.t())
        return (a + b).view(n, -1) / 2.0
    elif model == "rescal":  # rescal.py:15-53
        d = s_emb.size(-1)
        mix = p_emb.view(-1, d, d)
        if combine == "spo":
            out = (s_emb.unsqueeze(1).bmm(mix).view(n, d) * o_emb).sum(dim=-1)
        elif combine == "sp_":
            out = s_emb.unsqueeze(1).bmm(mix).view(n, d).mm(o_emb.t())
        else:
            out = mix.bmm(o_emb.unsqueeze(2)).view(n, d).mm(s_emb.t())
    elif model == "transe":  # transe.py:14-24
        if combine == "spo":
            out = -F.pairwise_distance(s_emb + p_emb, o_emb, p=l_norm)
        elif combine == "sp_":
            out = -torch.cdist(s_emb + p_emb, o_emb, p=l_norm)
        else:
            out = -torch.cdist(o_emb - p_emb, s_emb, p=l_norm)
    elif model == "rotate":  # rotate.py:18-63,97-145 (positive distance in this revision)
        s_re, s_im = torch.chunk(s_emb, 2, dim=1)
        o_re, o_im = torch.chunk(o_emb, 2, dim=1)
        p_re, p_im = torch.cos(p_emb), torch.sin(p_emb)
        if combine == "spo":
            sp_re = s_re * p_re - s_im * p_im
            sp_im = s_re * p_im + s_im * p_re
            out = torch.norm(_rot_abs(sp_re - o_re, sp_im - o_im), dim=1, p=l_norm)
        elif combine == "sp_":
            sp_re = s_re * p_re - s_im * p_im
            sp_im = s_re * p_im + s_im * p_re
            out = torch.norm(
                _rot_abs(sp_re.unsqueeze(1) - o_re, sp_im.unsqueeze(1) - o_im), dim=2, p=l_norm)
        else:
            sp_re = s_re.unsqueeze(1) * p_re - s_im.unsqueeze(1) * p_im  # s x p x dim
            sp_im = s_re.unsqueeze(1) * p_im + s_im.unsqueeze(1) * p_re
            out = torch.norm(_rot_abs(sp_re - o_re, sp_im - o_im), dim=2, p=l_norm).t()
    else:
        raise ValueError(model)
    return out.view(n, -1)


# --------------------------------------------------------------------------------------
# a4-a7: KgeModel.score_*  (kge_model.py:620-746), shared entity embedder
# --------------------------------------------------------------------------------------
def score_spo(model, ent_w, rel_w, s, p, o, l_norm=1.0):
    return score_emb(model, embed(ent_w, s), embed(rel_w, p), embed(ent_w, o), "spo", l_norm).view(-1)


def score_sp(model, ent_w, rel_w, s, p, o=None, l_norm=1.0):
    cand = embed_all(ent_w) if o is None else embed(ent_w, o)
    return score_emb(model, embed(ent_w, s), embed(rel_w, p), cand, "sp_", l_norm)


def score_po(model, ent_w, rel_w, p, o, s=None, l_norm=1.0):
    cand = embed_all(ent_w) if s is None else embed(ent_w, s)
    return score_emb(model, cand, embed(rel_w, p), embed(ent_w, o), "_po", l_norm)


def score_so(model, ent_w, rel_w, s, o, p=None, l_norm=1.0):
    cand = embed_all(rel_w) if p is None else embed(rel_w, p)
    return score_emb(model, embed(ent_w, s), cand, embed(ent_w, o), "s_o", l_norm)


def score_sp_po(model, ent_w, rel_w, s, p, o, entity_subset=None, l_norm=1.0):
    se, pe, oe = embed(ent_w, s), embed(rel_w, p), embed(ent_w, o)
    cand = embed_all(ent_w) if entity_subset is None else embed(ent_w, entity_subset)
    sp = score_emb(model, se, pe, cand, "sp_", l_norm)
    po = score_emb(model, cand, pe, oe, "_po", l_norm)
    return torch.cat((sp, po), dim=1)


def reciprocal_score_sp_po(model, ent_w, rel_w, num_rel, s, p, o, entity_subset=None, l_norm=1.0):
    """reciprocal_relations_model.py:79-106: both directions are "sp_", inverse uses p + R."""
    se, oe = embed(ent_w, s), embed(ent_w, o)
    pe, pinv = embed(rel_w, p), embed(rel_w, p + num_rel)
    cand = embed_all(ent_w) if entity_subset is None else embed(ent_w, entity_subset)
    sp = score_emb(model, se, pe, cand, "sp_", l_norm)
    po = score_emb(model, oe, pinv, cand, "sp_", l_norm)
    return torch.cat((sp, po), dim=1)


def reciprocal_score_spo(model, ent_w, rel_w, num_rel, s, p, o, direction, l_norm=1.0):
    """reciprocal_relations_model.py:56-63: "o" scores (s, p, o); "s" scores (o, p + R, s); no undirected form."""
    if direction == "o":
        return score_spo(model, ent_w, rel_w, s, p, o, l_norm)
    if direction == "s":
        return score_spo(model, ent_w, rel_w, o, p + num_rel, s, l_norm)
    raise Exception("The reciprocal relations model cannot compute undirected spo scores.")


def reciprocal_score_po(model, ent_w, rel_w, num_rel, p, o, s=None, l_norm=1.0):
    """reciprocal_relations_model.py:65-74: subjects ranked with the inverse relation, combine "sp_"."""
    cand = embed_all(ent_w) if s is None else embed(ent_w, s)
    return score_emb(model, embed(ent_w, o), embed(rel_w, p + num_rel), cand, "sp_", l_norm)


def batch_1vsall_reciprocal(model, prm: "Params", num_rel: int, triples: torch.Tensor, loss_name="kl", offset=0.0,
                            l_norm=1.0):
    """train.py:1032-1062 on a ReciprocalRelationsModel: the _po pass goes through reciprocal_score_po."""
    loss = make_loss(loss_name, offset)
    t = triples.long()
    b = len(t)
    l_sp = loss(score_sp(model, prm.ent, prm.rel, t[:, 0], t[:, 1], l_norm=l_norm), t[:, 2]) / b
    l_sp.backward()
    l_po = loss(reciprocal_score_po(model, prm.ent, prm.rel, num_rel, t[:, 1], t[:, 2], l_norm=l_norm), t[:, 0]) / b
    l_po.backward()
    return l_sp.item() + l_po.item()


# --------------------------------------------------------------------------------------
# a16/a17: losses (util/loss.py:137-159, 192-213).  Un-normalised sums; callers divide by B.
# --------------------------------------------------------------------------------------
def loss_kl(scores: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    if labels.dim() == 1:
        return F.cross_entropy(scores, labels.long(), reduction="sum")
    return F.kl_div(F.log_softmax(scores, dim=1), F.normalize(labels.float(), p=1, dim=1),
                    reduction="sum")


def loss_bce(scores: torch.Tensor, labels: torch.Tensor, offset: float = 0.0) -> torch.Tensor:
    if labels.dim() == 1:  # loss.py:105-117
        m = torch.zeros(scores.shape, dtype=torch.float)
        m[range(len(scores)), labels.long()] = 1.0
        labels = m
    if offset != 0.0:
        scores = scores + offset
    return F.binary_cross_entropy_with_logits(scores.view(-1), labels.view(-1), reduction="sum")


def make_loss(name: str, offset: float = 0.0):
    if name == "kl":
        return loss_kl
    if name == "bce":
        return lambda sc, lb: loss_bce(sc, lb, offset)
    raise ValueError(name)


# --------------------------------------------------------------------------------------
# a28: KvsAllIndex as CSR arrays  (indexing.py:36-55, 86-98)
# --------------------------------------------------------------------------------------
_KEYS = {"sp": ((0, 1), 2), "po": ((1, 2), 0), "so": ((0, 2), 1)}


def kvsall_index(triples: np.ndarray, key: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (keys [K,2], offsets [K+1], values [N]) for key in {"sp","po","so"}.

    Sort by value column, then stable-sort by the key columns from last to first; unique key
    rows give the CSR offsets.  Duplicate triples are kept (indexing.py:89-98).
    """
    key_cols, value_col = _KEYS[key]
    t = np.asarray(triples)
    t = t[np.argsort(t[:, value_col])]
    for c in key_cols[::-1]:
        t = t[np.argsort(t[:, c], kind="stable")]
    keys, first = np.unique(t[:, list(key_cols)], axis=0, return_index=True)
    offsets = np.append(first, len(t)).astype(np.int64)
    return keys, offsets, t[:, value_col].copy()


class CsrIndex:
    """dict-backed key lookup over the CSR arrays (indexing.py:44-71)."""

    def __init__(self, triples: np.ndarray, key: str):
        self.keys, self.offsets, self.values = kvsall_index(triples, key)
        self._row = {(int(a), int(b)): i for i, (a, b) in enumerate(self.keys)}

    def get(self, k) -> np.ndarray:
        i = self._row.get((int(k[0]), int(k[1])))
        if i is None:
            return np.zeros(0, dtype=self.values.dtype)
        return self.values[self.offsets[i]:self.offsets[i + 1]]

    def __len__(self):
        return len(self.keys)


def sp_po_coords(batch: np.ndarray, num_entities: int, sp: CsrIndex, po: CsrIndex) -> np.ndarray:
    """a24: job/util.py:5-38.  (row, o') for known (s,p,o'); (row, E+s') for known (s',p,o)."""
    out: List[Tuple[int, int]] = []
    for i, (s, p, o) in enumerate(np.asarray(batch).tolist()):
        out += [(i, int(x)) for x in sp.get((s, p))]
        out += [(i, int(x) + num_entities) for x in po.get((p, o))]
    return np.asarray(out, dtype=np.int64).reshape(-1, 2)


# --------------------------------------------------------------------------------------
# a18-a21: training batch bodies + optimizer step.  Parameters are leaf tensors; gradients
# accumulate across the passes of one batch exactly as in train.py:309-376.
# --------------------------------------------------------------------------------------
class Params:
    def __init__(self, ent_w: torch.Tensor, rel_w: torch.Tensor):
        self.ent = ent_w.detach().clone().float().requires_grad_(True)
        self.rel = rel_w.detach().clone().float().requires_grad_(True)

    def zero_grad(self):
        self.ent.grad = None
        self.rel.grad = None

    def grads(self):
        ge = self.ent.grad if self.ent.grad is not None else torch.zeros_like(self.ent)
        gr = self.rel.grad if self.rel.grad is not None else torch.zeros_like(self.rel)
        return ge.detach().clone(), gr.detach().clone()


def batch_1vsall(model, prm: Params, triples: torch.Tensor, loss_name="kl", offset=0.0, l_norm=1.0):
    """train.py:1032-1062: sp_ pass then _po pass, each loss / batch_size, grads accumulate."""
    loss = make_loss(loss_name, offset)
    t = triples.long()
    b = len(t)
    l_sp = loss(score_sp(model, prm.ent, prm.rel, t[:, 0], t[:, 1], l_norm=l_norm), t[:, 2]) / b
    l_sp.backward()
    l_po = loss(score_po(model, prm.ent, prm.rel, t[:, 1], t[:, 2], l_norm=l_norm), t[:, 0]) / b
    l_po.backward()
    return l_sp.item() + l_po.item()


def kvsall_collate(example_ids: Sequence[int], indexes: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]]):
    """train.py:590-677 for query types ["sp_", "_po"] (in that order).

    `indexes` = CSR triples (keys, offsets, values) per query type.  Returns queries [B,2],
    label_coords [nnz,2] (row, label), query_type [B].
    """
    ends = np.cumsum([len(ix[0]) for ix in indexes])
    queries, qtype, coords = [], [], []
    for row, ex in enumerate(example_ids):
        t = int(np.searchsorted(ends, ex, side="right"))
        local = ex - (0 if t == 0 else int(ends[t - 1]))
        keys, off, vals = indexes[t]
        queries.append(keys[local])
        qtype.append(t)
        for v in vals[off[local]:off[local + 1]]:
            coords.append((row, int(v)))
    return (np.asarray(queries, dtype=np.int64).reshape(-1, 2),
            np.asarray(coords, dtype=np.int64).reshape(-1, 2),
            np.asarray(qtype, dtype=np.int64))


def batch_kvsall(model, prm: Params, queries, label_coords, query_type, num_entities, num_relations,
                 loss_name="bce", offset=0.0, label_smoothing=0.0, l_norm=1.0):
    """train.py:679-756 for query types sp_ (0) and _po (1).

    Dense labels [B, max(E,R)] from coords; per query type: score, loss / (whole) batch size,
    backward.  The reported loss is the value of the last non-empty query type (train.py:747).
    """
    loss = make_loss(loss_name, offset)
    q = torch.as_tensor(queries).long()
    b = len(q)
    coords = torch.as_tensor(label_coords).long()
    labels = torch.zeros(b, max(num_entities, num_relations))
    labels.index_put_((coords[:, 0], coords[:, 1]), torch.ones(len(coords)), accumulate=True)
    qt = torch.as_tensor(query_type)
    reported = 0.0
    total = 0.0
    for t, name in enumerate(("sp_", "_po")):
        rows = (qt == t).nonzero().view(-1)
        if len(rows) == 0:
            continue
        lab = labels[rows, :num_entities]
        if label_smoothing > 0.0:
            lab = (1.0 - label_smoothing) * lab + 1.0 / lab.size(1)
        if name == "sp_":
            sc = score_sp(model, prm.ent, prm.rel, q[rows, 0], q[rows, 1], l_norm=l_norm)
        else:
            sc = score_po(model, prm.ent, prm.rel, q[rows, 0], q[rows, 1], l_norm=l_norm)
        lv = loss(sc, lab) / b
        reported = lv.item()
        total += reported
        lv.backward()
    return reported, total


def batch_negative_sampling(model, prm: Params, triples, negatives: Dict[int, torch.Tensor],
                            loss_name="kl", offset=0.0, l_norm=1.0):
    """train.py:823-999, implementation "triple": per slot with negatives, score the positive
    (column 0) and the N corrupted triples row-wise; labels[:,0]=1; loss / batch size."""
    loss = make_loss(loss_name, offset)
    t = triples.long()
    b = len(t)
    total = 0.0
    for slot in (S, P, O):
        neg = negatives.get(slot)
        if neg is None or neg.shape[1] == 0:
            continue
        n = neg.shape[1]
        rep = t.repeat(1, 1 + n).view(-1, 3)
        rep[:, slot] = torch.cat((t[:, [slot]], neg.long()), 1).view(-1)
        sc = score_spo(model, prm.ent, prm.rel, rep[:, 0], rep[:, 1], rep[:, 2], l_norm).view(b, -1)
        labels = torch.zeros(b, 1 + n)
        labels[:, 0] = 1
        lv = loss(sc, labels) / b
        total += lv.item()
        lv.backward()
    return total


def make_optimizer(name: str, prm: Params, **kw):
    """util/optimizer.py:10-17: getattr(torch.optim, name)(params, **args)."""
    return getattr(torch.optim, name)([prm.ent, prm.rel], **kw)


# --------------------------------------------------------------------------------------
# a23-a27: filtered entity ranking  (job/entity_ranking.py:112-217, 428-577; eval.py:138-171)
# --------------------------------------------------------------------------------------
def ranks_and_ties(scores: torch.Tensor, true_scores: torch.Tensor):
    """entity_ranking.py:507-529."""
    sc = scores.clone()
    sc[torch.isnan(sc)] = float("-Inf")
    ts = true_scores.clone()
    ts[torch.isnan(ts)] = float("-Inf")
    rank = torch.sum(sc > ts.view(-1, 1), dim=1, dtype=torch.long)
    ties = torch.sum(sc == ts.view(-1, 1), dim=1, dtype=torch.long)
    return rank, ties


def final_ranks(rank, ties, tie_handling="rounded_mean_rank"):
    """entity_ranking.py:531-551 (0-based ranks)."""
    if tie_handling == "rounded_mean_rank":
        return rank + ties // 2
    if tie_handling == "best_rank":
        return rank
    if tie_handling == "worst_rank":
        return rank + ties - 1
    raise NotImplementedError(tie_handling)


def _dense_labels(coords: torch.Tensor, b: int, e: int, lo: int, hi: int) -> torch.Tensor:
    """entity_ranking.py:122-135 + 428-467: +inf at known answers, sliced to [lo,hi) twice."""
    lab = torch.zeros(b, 2 * (hi - lo))
    if len(coords):
        r, c = coords[:, 0], coords[:, 1]
        m_sp = (c >= lo) & (c < hi)
        m_po = (c >= lo + e) & (c < hi + e)
        lab[r[m_sp], c[m_sp] - lo] = float("inf")
        lab[r[m_po], c[m_po] - e - lo + (hi - lo)] = float("inf")
    return lab


def entity_ranking_batch(model, ent_w, rel_w, batch, coords, test_coords=None, chunk_size=-1,
                         l_norm=1.0, score_fn=None):
    """One batch of EntityRankingJob._run (entity_ranking.py:112-217).

    Returns dict name -> (rank, ties) for names s_raw,o_raw,s_filt,o_filt[,s_filt_test,o_filt_test].
    `score_fn(s,p,o,subset)` overrides model.score_sp_po (used for reciprocal models).
    """
    with torch.no_grad():
        bt = torch.as_tensor(batch)
        s, p, o = bt[:, 0], bt[:, 1], bt[:, 2]
        b, e = len(bt), ent_w.shape[0]
        coords = torch.as_tensor(coords).long().view(-1, 2)
        rankings = [("_raw", None), ("_filt", coords)]
        if test_coords is not None:
            rankings.append(("_filt_test", torch.as_tensor(test_coords).long().view(-1, 2)))
        o_true = score_spo(model, ent_w, rel_w, s, p, o, l_norm)
        s_true = score_spo(model, ent_w, rel_w, s, p, o, l_norm)
        acc = {d + n: [torch.zeros(b, dtype=torch.long), torch.zeros(b, dtype=torch.long)]
               for n, _ in rankings for d in "so"}
        cs = chunk_size if chunk_size > -1 else e
        for c in range(math.ceil(e / cs)):
            lo, hi = cs * c, min(cs * (c + 1), e)
            subset = torch.arange(lo, hi)
            if score_fn is None:
                scores = score_sp_po(model, ent_w, rel_w, s, p, o, subset, l_norm)
            else:
                scores = score_fn(s, p, o, subset)
            sc_sp, sc_po = scores[:, :hi - lo], scores[:, hi - lo:]
            s_in = (lo <= s) & (s < hi)
            o_in = (lo <= o) & (o < hi)
            o_loc = (o[o_in] - lo).long()
            s_loc = (s[s_in] - lo).long()
            sc_sp[o_in, o_loc] = o_true[o_in]
            sc_po[s_in, s_loc] = s_true[s_in]
            for name, cc in rankings:
                if cc is not None:
                    lab = _dense_labels(cc, b, e, lo, hi)
                    lab[o_in, o_loc] = 0
                    lab[s_in, s_loc + (hi - lo)] = 0
                    sc_sp = sc_sp - lab[:, :hi - lo]
                    sc_po = sc_po - lab[:, hi - lo:]
                o_r, o_t = ranks_and_ties(sc_sp, o_true)
                s_r, s_t = ranks_and_ties(sc_po, s_true)
                acc["s" + name][0] += s_r
                acc["s" + name][1] += s_t
                acc["o" + name][0] += o_r
                acc["o" + name][1] += o_t
        return {k: (v[0], v[1]) for k, v in acc.items()}


def rank_histogram(ranks: torch.Tensor, num_entities: int) -> torch.Tensor:
    """eval.py:138-171 (hist_all): float32 [E] histogram of 0-based ranks."""
    h = torch.zeros(num_entities, dtype=torch.float)
    for r in ranks.tolist():
        h[r] += 1
    return h


def metrics_from_hist(hist: torch.Tensor, hits_at_k=(1, 3, 10)) -> Dict[str, float]:
    """entity_ranking.py:553-577."""
    n = torch.sum(hist).item()
    e = len(hist)
    ks = [min(k, e) for k in hits_at_k]  # eval.py:19-24
    if n <= 0:
        return {"mean_rank": 0.0, "mean_reciprocal_rank": 0.0, **{f"hits_at_{k}": 0.0 for k in ks}}
    ranks = torch.arange(1, e + 1).float()
    out = {"mean_rank": torch.sum(hist * ranks).item() / n,
           "mean_reciprocal_rank": torch.sum(hist * (1.0 / ranks)).item() / n}
    cum = (torch.cumsum(hist[:max(ks)], dim=0) / n).tolist()
    for k in ks:
        out[f"hits_at_{k}"] = cum[k - 1]
    return out


def entity_ranking(model, ent_w, rel_w, triples, filter_triples: Sequence[np.ndarray],
                   test_triples: Optional[np.ndarray] = None, batch_size=64, chunk_size=-1,
                   tie_handling="rounded_mean_rank", hits_at_k=(1, 3, 10), l_norm=1.0):
    """EntityRankingJob._run over `triples` (entity_ranking.py:79-426), metrics for _raw, _filt
    (and _filt_test).  `filter_triples` = one [N,3] array per split in entity_ranking.filter_splits."""
    e = ent_w.shape[0]
    fidx = [(CsrIndex(t, "sp"), CsrIndex(t, "po")) for t in filter_triples]
    tidx = (CsrIndex(test_triples, "sp"), CsrIndex(test_triples, "po")) if test_triples is not None else None
    names = ["_raw", "_filt"] + (["_filt_test"] if tidx else [])
    hists = {n: torch.zeros(e) for n in names}
    all_ranks = {d + n: [] for n in names for d in "so"}
    tr = np.asarray(triples)
    for lo in range(0, len(tr), batch_size):
        bt = tr[lo:lo + batch_size]
        coords = np.concatenate([sp_po_coords(bt, e, a, b) for a, b in fidx]) if fidx else np.zeros((0, 2), np.int64)
        tcoords = sp_po_coords(bt, e, *tidx) if tidx else None
        res = entity_ranking_batch(model, ent_w, rel_w, bt, coords, tcoords, chunk_size, l_norm)
        for n in names:
            for d in "so":
                r = final_ranks(*res[d + n], tie_handling)
                all_ranks[d + n].append(r)
                hists[n] += rank_histogram(r, e)
    metrics = {n: metrics_from_hist(h, hits_at_k) for n, h in hists.items()}
    return metrics, {k: torch.cat(v) if v else torch.zeros(0, dtype=torch.long) for k, v in all_ranks.items()}


# ---------------------------------------------------------------------------------------------
# rows around the path
# ---------------------------------------------------------------------------------------------
def lp_penalty(weight: torch.Tensor, p: int, reg_weight: float, indexes: Optional[torch.Tensor] = None,
               weighted: bool = False) -> torch.Tensor:
    """lookup_embedder.py:112-158: unweighted `w/p * ||W||_p^p` over the whole table (:126-134), or the weighted
    per-distinct-index form divided by len(indexes) (:135-155).  Differentiable w.r.t. `weight`."""
    if not weighted:
        return (reg_weight / p * weight.norm(p=p) ** p).sum()
    unique, counts = torch.unique(indexes, return_counts=True)
    prm = F.embedding(unique.long(), weight)
    if p % 2 == 1:
        prm = torch.abs(prm)
    return (reg_weight / p * (prm ** p * counts.float().view(-1, 1))).sum() / len(indexes)


def model_penalties(ent_w: torch.Tensor, rel_w: torch.Tensor, triples: torch.Tensor, p: int, w_ent: float, w_rel: float,
                    weighted: bool) -> List[torch.Tensor]:
    """kge_model.py:588-606: s-embedder, p-embedder, o-embedder terms in that order (the shared entity table is
    penalised once per slot)."""
    return [lp_penalty(ent_w, p, w_ent, triples[:, S], weighted), lp_penalty(rel_w, p, w_rel, triples[:, P], weighted),
            lp_penalty(ent_w, p, w_ent, triples[:, O], weighted)]


def frequency_histograms(ranks: Dict[str, torch.Tensor], triples: torch.Tensor, num_entities: int,
                         percentiles: Dict[str, Dict[str, Sequence[int]]]) -> Dict[str, torch.Tensor]:
    """eval.py:201-224 (hist_per_frequency_percentile) for one filter setting: "subject_<perc>" counts the subject ranks
    of triples whose subject is in the percentile, "object_<perc>" the object ranks likewise, "relation_<perc>" both."""
    h: Dict[str, torch.Tensor] = {}
    for arg, percs in percentiles.items():
        for perc in percs:
            h[f"{arg}_{perc}"] = torch.zeros(num_entities)
    s, p, o = (triples[:, i].tolist() for i in range(3))
    for perc in percentiles["subject"]:
        subj, rel, obj = (set(int(x) for x in percentiles[a][perc]) for a in ("subject", "relation", "object"))
        for i, r in enumerate(ranks["s"].tolist()):
            if s[i] in subj:
                h[f"subject_{perc}"][r] += 1
            if p[i] in rel:
                h[f"relation_{perc}"][r] += 1
        for i, r in enumerate(ranks["o"].tolist()):
            if o[i] in obj:
                h[f"object_{perc}"][r] += 1
            if p[i] in rel:
                h[f"relation_{perc}"][r] += 1
    return h


def grouped_rank_histograms(ranks: Dict[str, torch.Tensor], relations: torch.Tensor, num_entities: int,
                            relations_per_type: Optional[Dict[str, Sequence[int]]] = None,
                            head_and_tail: bool = False) -> Dict[str, torch.Tensor]:
    """eval.py:138-198 for one filter setting: ranks = {"s": subject ranks, "o": object ranks} of the same triples;
    returns histograms keyed "all", "head", "tail", "<type>", "<type>_head", "<type>_tail"."""
    h: Dict[str, torch.Tensor] = {"all": torch.zeros(num_entities)}

    def add(key, r):
        h.setdefault(key, torch.zeros(num_entities))
        h[key][r] += 1

    p = relations.tolist()
    for side, name in (("o", "tail"), ("s", "head")):
        for i, r in enumerate(ranks[side].tolist()):
            add("all", r)
            if head_and_tail:
                add(name, r)
            for t, rels in (relations_per_type or {}).items():
                if p[i] in set(int(x) for x in rels):
                    add(t, r)
                    if head_and_tail:
                        add(f"{t}_{name}", r)
    return h
