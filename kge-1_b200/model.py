"""Host-side mirror of the reference's model API for the hot path.

Same names, argument meaning and error behaviour as kge/model/kge_model.py and
kge/model/embedder/lookup_embedder.py so that the reference's jobs (TrainingJob1vsAll / KvsAll /
NegativeSampling, EntityRankingJob) can call these classes unchanged (see INTEGRATION.md for the
registration stub).  All arithmetic runs in the CUDA kernels behind include/kgeb200.h.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import lib, ops

SCORER_MODELS = ("distmult", "complex", "cp", "simple", "rescal", "transe", "rotate")


class LookupEmbedder(torch.nn.Module):
    """kge/model/embedder/lookup_embedder.py:14-158.

    The table is a torch.nn.Embedding so that `state_dict()` keeps the reference's checkpoint key
    `_embeddings.weight`; it is only used as parameter storage -- gathers run in kgeb_gather_rows and
    `embed_all()` returns the weight in place (the reference's arange-gather copy, lookup_embedder.py:102-107,
    is eliminated; kernels read the table where it lies).
    """

    def __init__(self, vocab_size: int, dim: int, sparse: bool = False, dropout: float = 0.0,
                 normalize_p: float = -1.0, regularize: str = "lp", regularize_weight: float = 0.0,
                 regularize_p: int = 2, regularize_weighted: bool = False, initialize: str = "normal_",
                 initialize_args: Optional[dict] = None, configuration_key: str = "lookup_embedder"):
        super().__init__()
        if regularize not in ("", "lp"):
            raise ValueError(f"Invalid value regularize={regularize}")
        self.configuration_key = configuration_key
        self.vocab_size, self.dim = int(vocab_size), int(dim)
        self.sparse = bool(sparse)
        self.normalize_p = normalize_p
        self.regularize, self.regularize_weight = regularize, float(regularize_weight)
        self.regularize_p, self.regularize_weighted = regularize_p, bool(regularize_weighted)
        self._embeddings = torch.nn.Embedding(self.vocab_size, self.dim, sparse=self.sparse)
        args = {"mean": 0.0, "std": 0.1} if initialize_args is None and initialize == "normal_" else (initialize_args or {})
        getattr(torch.nn.init, initialize)(self._embeddings.weight.data, **args)
        self.dropout = torch.nn.Dropout(max(float(dropout), 0.0))

    # -- kge_model.py:333-339 -----------------------------------------------------------------
    def embed(self, indexes: Tensor) -> Tensor:
        return self._postprocess(ops.gather_rows(self._embeddings.weight, indexes, self.sparse))

    def embed_all(self) -> Tensor:
        return self._postprocess(self._embeddings.weight)

    def _postprocess(self, embeddings: Tensor) -> Tensor:
        if self.dropout.p > 0:
            embeddings = self.dropout(embeddings)
        return embeddings

    @property
    def weight(self) -> Tensor:
        return self._embeddings.weight

    def prepare_job(self, job, **kwargs):
        """lookup_embedder.py:58-75: per-batch renormalisation hook (replaces the weight tensor)."""
        if self.normalize_p > 0:
            def normalize_embeddings(job):
                with torch.no_grad():
                    self._embeddings.weight = torch.nn.Parameter(
                        torch.nn.functional.normalize(self._embeddings.weight, p=self.normalize_p, dim=-1))
            job.pre_batch_hooks.append(normalize_embeddings)

    def penalty(self, **kwargs) -> List[Tuple[str, Tensor]]:
        """lookup_embedder.py:112-158 (Lp penalty, weighted or unweighted)."""
        if self.regularize == "" or self.regularize_weight == 0.0:
            return []
        p, w = self.regularize_p, self.regularize_weight
        if not self.regularize_weighted:
            value = ops.lp_penalty(self._embeddings.weight, p, w)
        else:
            value = ops.lp_penalty(self._embeddings.weight, p, w, kwargs["indexes"])
        return [(f"{self.configuration_key}.L{p}_penalty", value)]


class RelationalScorer(torch.nn.Module):
    """kge_model.py:91-182.  One class serves the seven scorers; `model` selects the formulas."""

    def __init__(self, model: str, l_norm: float = 1.0, math_mode: int = lib.MATH_FP32):
        super().__init__()
        if model not in SCORER_MODELS:
            raise ValueError(f"unknown scorer {model}")
        self.model, self._norm = model, l_norm
        self.math_mode = math_mode  # arithmetic of the all-entity DOT tiles (fp32 FFMA | TF32 tcgen05)
        self.kind = ops.pair_kind(model, l_norm)

    def score_emb_spo(self, s_emb: Tensor, p_emb: Tensor, o_emb: Tensor) -> Tensor:
        return self.score_emb(s_emb, p_emb, o_emb, "spo")

    def score_emb(self, s_emb: Tensor, p_emb: Tensor, o_emb: Tensor, combine: str) -> Tensor:
        n = p_emb.size(0)
        if combine == "spo":
            if s_emb.size(0) != n or o_emb.size(0) != n:
                raise ValueError("spo scoring needs the same number of s, p and o embeddings")
            out = ops.score_spo_emb(self.model, self._norm, s_emb, p_emb, o_emb)
        elif combine == "sp_":
            q = ops.query_build(self.model, lib.SP_, s_emb, p_emb)
            out = ops.score_all(self.kind, q, o_emb, self._math(o_emb))
        elif combine == "_po":
            q = ops.query_build(self.model, lib._PO, o_emb, p_emb)
            out = ops.score_all(self.kind, q, s_emb, self._math(s_emb))
        elif combine == "s_o":
            # generic fallback (kge_model.py:171-178): relation-wide, tiny; routed through the spo kernel
            n = s_emb.size(0)
            n_p = p_emb.size(0)
            out = ops.score_spo_emb(self.model, self._norm, s_emb.repeat_interleave(n_p, 0), p_emb.repeat((n, 1)),
                                    o_emb.repeat_interleave(n_p, 0))
            if self.model in ("rescal", "transe"):
                # rescal.py:50-53 / transe.py:22-24 re-view the result with n = p_emb.size(0)
                return out.view(n, -1).view(n_p, -1)
        else:
            raise ValueError('cannot handle combine="{}"'.format(combine))
        return out.view(n, -1)

    def _math(self, cand: Tensor) -> int:
        if self.kind != lib.DOT or self.math_mode != lib.MATH_TF32:
            return lib.MATH_FP32
        return lib.MATH_TF32 if (cand.shape[1] % 4 == 0 and cand.shape[1] <= 256) else lib.MATH_FP32


class KgeModel(torch.nn.Module):
    """kge_model.py:342-746: entity + relation LookupEmbedders and a RelationalScorer."""

    def __init__(self, model: str, num_entities: int, num_relations: int, dim: int, l_norm: float = 1.0,
                 relation_dim: int = -1, math_mode: int = lib.MATH_FP32, entity_embedder: Optional[dict] = None,
                 relation_embedder: Optional[dict] = None):
        super().__init__()
        if model in ("complex", "cp", "simple", "rotate") and dim % 2 != 0:
            names = {"complex": "ComplEx", "cp": "CP", "simple": "SimplE", "rotate": "RotatE"}
            raise ValueError(f"{names[model]} requires embeddings of even dimensionality (got {dim})")
        self.model = model
        self.num_entities, self.num_relations = num_entities, num_relations
        rdim = relation_dim if relation_dim > 0 else ops.relation_dim(model, dim)
        rel_args = dict(relation_embedder or {})
        if model == "rotate" and "initialize" not in rel_args:  # rotate.yaml:12-26 phases U(-pi, pi)
            rel_args.update(initialize="uniform_", initialize_args={"a": -math.pi, "b": math.pi})
        self._entity_embedder = LookupEmbedder(num_entities, dim, configuration_key=f"{model}.entity_embedder",
                                               **(entity_embedder or {}))
        self._relation_embedder = LookupEmbedder(num_relations, rdim, configuration_key=f"{model}.relation_embedder",
                                                 **rel_args)
        self._scorer = RelationalScorer(model, l_norm, math_mode)

    @staticmethod
    def create(model: str, num_entities: int, num_relations: int, dim: int, **kw) -> "KgeModel":
        if model == "reciprocal_relations_model":
            return ReciprocalRelationsModel(kw.pop("base_model"), num_entities, num_relations, dim, **kw)
        return KgeModel(model, num_entities, num_relations, dim, **kw)

    # -- getters (kge_model.py:608-618) ------------------------------------------------------
    def get_s_embedder(self): return self._entity_embedder
    def get_o_embedder(self): return self._entity_embedder
    def get_p_embedder(self): return self._relation_embedder
    def get_scorer(self): return self._scorer

    def prepare_job(self, job, **kwargs):
        self._entity_embedder.prepare_job(job, **kwargs)
        self._relation_embedder.prepare_job(job, **kwargs)

    def penalty(self, **kwargs):
        """kge_model.py:588-606: entity penalties see the s and o columns, relation penalties p."""
        if "batch" in kwargs and "triples" in kwargs["batch"]:
            t = kwargs["batch"]["triples"].to(self._entity_embedder.weight.device)
            return (self._entity_embedder.penalty(indexes=t[:, 0], **kwargs)
                    + self._relation_embedder.penalty(indexes=t[:, 1], **kwargs)
                    + self._entity_embedder.penalty(indexes=t[:, 2], **kwargs))
        return (self._entity_embedder.penalty(**kwargs) + self._relation_embedder.penalty(**kwargs)
                + self._entity_embedder.penalty(**kwargs))

    # -- scoring (kge_model.py:620-746) -------------------------------------------------------
    def _no_dropout(self) -> bool:
        return not self.training or (self._entity_embedder.dropout.p == 0 and self._relation_embedder.dropout.p == 0)

    def score_spo(self, s: Tensor, p: Tensor, o: Tensor, direction=None) -> Tensor:
        if self._no_dropout():  # fused gather + score
            return ops.score_spo_indexed(self.model, self._scorer._norm, self._entity_embedder.weight,
                                         self._relation_embedder.weight, s, p, o).view(-1)
        s_emb = self.get_s_embedder().embed(s)
        p_emb = self.get_p_embedder().embed(p)
        o_emb = self.get_o_embedder().embed(o)
        return self._scorer.score_emb(s_emb, p_emb, o_emb, combine="spo").view(-1)

    def score_sp(self, s: Tensor, p: Tensor, o: Tensor = None) -> Tensor:
        s_emb = self.get_s_embedder().embed(s)
        p_emb = self.get_p_embedder().embed(p)
        o_emb = self.get_o_embedder().embed_all() if o is None else self.get_o_embedder().embed(o)
        return self._scorer.score_emb(s_emb, p_emb, o_emb, combine="sp_")

    def score_po(self, p: Tensor, o: Tensor, s: Tensor = None) -> Tensor:
        s_emb = self.get_s_embedder().embed_all() if s is None else self.get_s_embedder().embed(s)
        o_emb = self.get_o_embedder().embed(o)
        p_emb = self.get_p_embedder().embed(p)
        return self._scorer.score_emb(s_emb, p_emb, o_emb, combine="_po")

    def score_so(self, s: Tensor, o: Tensor, p: Tensor = None) -> Tensor:
        s_emb = self.get_s_embedder().embed(s)
        o_emb = self.get_o_embedder().embed(o)
        p_emb = self.get_p_embedder().embed_all() if p is None else self.get_p_embedder().embed(p)
        return self._scorer.score_emb(s_emb, p_emb, o_emb, combine="s_o")

    def score_sp_po(self, s: Tensor, p: Tensor, o: Tensor, entity_subset: Tensor = None) -> Tensor:
        s_emb = self.get_s_embedder().embed(s)
        p_emb = self.get_p_embedder().embed(p)
        o_emb = self.get_o_embedder().embed(o)
        if entity_subset is not None:
            all_entities = self.get_s_embedder().embed(entity_subset)
        else:
            all_entities = self.get_s_embedder().embed_all()
        sp_scores = self._scorer.score_emb(s_emb, p_emb, all_entities, combine="sp_")
        po_scores = self._scorer.score_emb(all_entities, p_emb, o_emb, combine="_po")
        return torch.cat((sp_scores, po_scores), dim=1)

    # -- query vectors for the fused kernels (not in the reference API) ---------------------------
    def queries(self, combine: int, a: Tensor, p: Tensor) -> Tensor:
        """Q[i,:] = q(a_i, p_i) with autograd through the gathers."""
        a_emb = self._entity_embedder.embed(a)
        p_emb = self._relation_embedder.embed(p)
        return ops.query_build(self.model, combine, a_emb, p_emb)


class ReciprocalRelationsModel(KgeModel):
    """reciprocal_relations_model.py:7-106: base model with 2R relations; _po becomes sp_ with p + R."""

    def __init__(self, base_model: str, num_entities: int, num_relations: int, dim: int, **kw):
        super().__init__(base_model, num_entities, 2 * num_relations, dim, **kw)
        self.num_relations = num_relations

    def score_spo(self, s, p, o, direction=None):
        if direction == "o":
            return super().score_spo(s, p, o, "o")
        elif direction == "s":
            return super().score_spo(o, p + self.num_relations, s, "o")
        raise Exception("The reciprocal relations model cannot compute undirected spo scores.")

    def score_po(self, p, o, s=None):
        s_emb = self.get_s_embedder().embed_all() if s is None else self.get_s_embedder().embed(s)
        p_emb = self.get_p_embedder().embed(p + self.num_relations)
        o_emb = self.get_o_embedder().embed(o)
        return self._scorer.score_emb(o_emb, p_emb, s_emb, combine="sp_")

    def score_so(self, s, o, p=None):
        raise Exception("The reciprocal relations model cannot score relations.")

    def score_sp_po(self, s, p, o, entity_subset=None):
        s_emb = self.get_s_embedder().embed(s)
        p_inv = self.get_p_embedder().embed(p + self.num_relations)
        p_emb = self.get_p_embedder().embed(p)
        o_emb = self.get_o_embedder().embed(o)
        if entity_subset is not None:
            all_entities = self.get_s_embedder().embed(entity_subset)
        else:
            all_entities = self.get_s_embedder().embed_all()
        sp_scores = self._scorer.score_emb(s_emb, p_emb, all_entities, combine="sp_")
        po_scores = self._scorer.score_emb(o_emb, p_inv, all_entities, combine="sp_")
        return torch.cat((sp_scores, po_scores), dim=1)
