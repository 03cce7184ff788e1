"""Rank histograms and ranking metrics on the device (kge/job/eval.py:138-224, entity_ranking.py:553-577) through
csrc/metrics.cu: no Python loop per rank, no per-id list membership tests, one small read-back per metric set."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import lib


def rank_hist(ranks: torch.Tensor, num_entities: int, hist: Optional[torch.Tensor] = None,
              mask: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hist[r] += 1 for every rank r (0-based, int64 CUDA) whose mask entry is set; float32 [num_entities] as
    eval.py:130-136 creates it."""
    ranks = lib_i64(ranks, "ranks")
    if hist is None:
        hist = torch.zeros(num_entities, dtype=torch.float32, device=ranks.device)
    lib.f32(hist, "hist")
    mptr = None
    if mask is not None:
        if not mask.is_cuda or mask.dtype != torch.uint8 or mask.numel() != ranks.numel() or not mask.is_contiguous():
            raise ValueError("mask must be a contiguous uint8 CUDA tensor with one entry per rank")
        mptr = mask.data_ptr()
    lib.call("kgeb_rank_hist", ranks.data_ptr(), mptr, ranks.numel(), num_entities, hist.data_ptr(),
             None if status is None else status.data_ptr(), lib.stream_ptr(ranks))
    return hist


def lib_i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU path exists)")
    return t.long().contiguous()


def isin_sorted(values: torch.Tensor, sorted_set: torch.Tensor) -> torch.Tensor:
    """uint8 mask of `values[i] in set` for an ascending int64 CUDA `sorted_set` (eval.py:187, 205-221)."""
    values = values.contiguous()
    vptr, v64 = lib.idx(values, "values")
    sorted_set = lib_i64(sorted_set, "sorted_set")
    mask = torch.empty(values.numel(), dtype=torch.uint8, device=values.device)
    lib.call("kgeb_isin_sorted", vptr, v64, values.numel(), sorted_set.data_ptr(), sorted_set.numel(), mask.data_ptr(),
             lib.stream_ptr(values))
    return mask


def rank_metrics(hist: torch.Tensor, hits_at_k_s: Sequence[int], suffix: str = "") -> Dict[str, float]:
    """EntityRankingJob._compute_metrics (entity_ranking.py:553-577): mean_rank, mean_reciprocal_rank, hits_at_k."""
    import ctypes
    lib.f32(hist, "hist")
    ks = list(hits_at_k_s)
    L = lib.load()
    ws = torch.empty(L.kgeb_rank_metrics_workspace_bytes(), dtype=torch.uint8, device=hist.device)
    out = torch.empty(3 + len(ks), dtype=torch.float64, device=hist.device)
    arr = (ctypes.c_int32 * max(len(ks), 1))(*ks)
    lib.call("kgeb_rank_metrics", hist.data_ptr(), hist.numel(), arr, len(ks), out.data_ptr(), ws.data_ptr(), ws.numel(),
             lib.stream_ptr(hist))
    vals = out.tolist()
    metrics = {"mean_rank" + suffix: vals[1], "mean_reciprocal_rank" + suffix: vals[2]}
    for k, v in zip(ks, vals[3:]):
        metrics["hits_at_{}{}".format(k, suffix)] = v
    return metrics
