"""torch.autograd bindings of the CUDA kernels (the host-side half of the drop-in boundary).

Each Function mirrors one ATen op sequence of the reference and calls
the C-ABI entry point that replaces it; backward passes call the matching *_bwd entry points and the
deterministic sort-based scatter.
"""
from __future__ import annotations

from typing import Optional

import weakref

import torch

from . import lib

_ws_cache = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only per-device scratch buffer (reused between calls on the same stream)."""
    key = (device.type, device.index)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


_mirror_cache = {}


def bf16_mirror(table: torch.Tensor) -> torch.Tensor:
    """bf16 copy of an fp32 table for KGEB_MATH_BF16, refreshed when the tensor's version counter moves
    (optimizer steps, renormalisation hooks and checkpoint loads all bump it).  An entry belongs to a STORAGE (weak
    reference: PyTorch keeps one Python object per live storage), not to an address: the caching allocator hands the
    address of a freed table to the next one of the same shape, and version counters of fresh tensors coincide."""
    t = table.detach()
    storage = t.untyped_storage()
    key = (t.data_ptr(), tuple(t.shape))
    hit = _mirror_cache.get(key)
    if hit is not None and hit[2]() is not storage:
        hit = None
    if hit is not None and hit[0] == table._version:
        return hit[1]
    buf = hit[1] if hit is not None else torch.empty(t.shape, dtype=torch.bfloat16, device=t.device)
    lib.call("kgeb_to_bf16", lib.f32(t.contiguous(), "table"), buf.data_ptr(), t.numel(), lib.stream_ptr(t))
    if len(_mirror_cache) > 16:
        _mirror_cache.clear()
    _mirror_cache[key] = (table._version, buf, weakref.ref(storage))
    return buf


def pair_kind(model: str, l_norm: float = 1.0) -> int:
    """Pair-score family of a scorer for its all-entity forms (include/kgeb200.h)."""
    if model in ("distmult", "complex", "cp", "simple", "rescal"):
        return lib.DOT
    if l_norm not in (1, 1.0, 2, 2.0):
        raise NotImplementedError(f"l_norm={l_norm}: only l_norm 1 and 2 are built")
    if model == "transe":
        return lib.NEG_L1 if float(l_norm) == 1.0 else lib.NEG_L2
    if model == "rotate":
        return lib.ROT_L1 if float(l_norm) == 1.0 else lib.ROT_L2
    raise ValueError(f"unknown model {model}")


def relation_dim(model: str, d: int) -> int:
    if model in ("cp", "rotate"):
        return d // 2
    if model == "rescal":
        return d * d
    return d


# ---------------------------------------------------------------------------------------------
# scatter (autograd of the gathers)
# ---------------------------------------------------------------------------------------------
def scatter_add_rows_(dense: torch.Tensor, indexes: torch.Tensor, rows: torch.Tensor,
                      perm: torch.Tensor = None) -> torch.Tensor:
    """dense[indexes[i], :] += rows[i, :] (sorted, deterministic).  perm (int32, optional) = stable argsort of indexes,
    e.g. from the batch collate: spares the device sort, same sums in the same order."""
    n = indexes.numel()
    if n == 0:
        return dense
    indexes = indexes.reshape(-1).contiguous()
    rows = rows.reshape(n, -1).contiguous()
    nbytes = lib.load().kgeb_scatter_workspace_bytes(n, rows.shape[1])
    ws = _workspace(dense.device, nbytes)
    ip, i64 = lib.idx(indexes)
    if perm is not None:
        if perm.dtype != torch.int32 or perm.numel() != n or not perm.is_cuda:
            raise ValueError("perm must be an int32 CUDA tensor with one entry per index")
        lib.call("kgeb_scatter_add_rows_perm", ip, i64, perm.contiguous().data_ptr(), lib.f32(rows, "rows"), n,
                 rows.shape[1], lib.f32(dense, "dense"), dense.shape[0], ws.data_ptr(), ws.numel(), lib.stream_ptr(dense))
        return dense
    lib.call("kgeb_scatter_add_rows", ip, i64, lib.f32(rows, "rows"), n, rows.shape[1], lib.f32(dense, "dense"),
             dense.shape[0], ws.data_ptr(), ws.numel(), lib.stream_ptr(dense))
    return dense


def segment_reduce_rows(indexes: torch.Tensor, rows: torch.Tensor):
    """Returns (distinct ids ascending [n], summed rows [n,d], count tensor [1]); only the first count rows are valid."""
    n = indexes.numel()
    indexes = indexes.reshape(-1).contiguous()
    rows = rows.reshape(n, -1).contiguous()
    d = rows.shape[1]
    ids = torch.empty(max(n, 1), dtype=torch.int64, device=rows.device)
    out = torch.empty(max(n, 1), d, dtype=torch.float32, device=rows.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=rows.device)
    ws = _workspace(rows.device, lib.load().kgeb_scatter_workspace_bytes(n, d))
    ip, i64 = lib.idx(indexes)
    lib.call("kgeb_segment_reduce_rows", ip, i64, lib.f32(rows, "rows"), n, d, ids.data_ptr(), out.data_ptr(),
             cnt.data_ptr(), ws.data_ptr(), ws.numel(), lib.stream_ptr(rows))
    return ids, out, cnt


# ---------------------------------------------------------------------------------------------
# K1 / K3: embedding gather
# ---------------------------------------------------------------------------------------------
class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, indexes, sparse):
        w = weight.detach()
        ind = indexes.reshape(-1).contiguous()
        out = torch.empty(ind.numel(), w.shape[1], dtype=torch.float32, device=w.device)
        ip, i64 = lib.idx(ind)
        lib.call("kgeb_gather_rows", lib.f32(w, "weight"), w.shape[0], w.shape[1], ip, i64, ind.numel(),
                 out.data_ptr(), lib.stream_ptr(w))
        ctx.save_for_backward(ind)
        ctx.shape = tuple(w.shape)
        ctx.sparse = sparse
        return out.view(*indexes.shape, w.shape[1])

    @staticmethod
    def backward(ctx, grad):
        (ind,) = ctx.saved_tensors
        grad = grad.reshape(ind.numel(), -1).contiguous()
        if ctx.sparse:
            ids, rows, cnt = segment_reduce_rows(ind, grad)
            k = int(cnt.item())
            g = torch.sparse_coo_tensor(ids[:k].unsqueeze(0), rows[:k], ctx.shape, device=grad.device)
            return g, None, None
        dense = torch.zeros(ctx.shape, dtype=torch.float32, device=grad.device)
        scatter_add_rows_(dense, ind, grad)
        return dense, None, None


def gather_rows(weight: torch.Tensor, indexes: torch.Tensor, sparse: bool = False) -> torch.Tensor:
    return _Gather.apply(weight, indexes, sparse)


# ---------------------------------------------------------------------------------------------
# K5: row-wise "spo" scoring on embedding matrices
# ---------------------------------------------------------------------------------------------
class _ScoreSpoEmb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model_id, l_norm, s_emb, p_emb, o_emb):
        s, p, o = (t.detach().contiguous() for t in (s_emb, p_emb, o_emb))
        n, d = s.shape
        out = torch.empty(n, dtype=torch.float32, device=s.device)
        lib.call("kgeb_score_spo", model_id, l_norm, lib.f32(s, "s_emb"), None, lib.f32(p, "p_emb"), None,
                 lib.f32(o, "o_emb"), None, 0, n, d, out.data_ptr(), lib.stream_ptr(s))
        ctx.save_for_backward(s, p, o)
        ctx.model_id, ctx.l_norm = model_id, l_norm
        return out

    @staticmethod
    def backward(ctx, gout):
        s, p, o = ctx.saved_tensors
        n, d = s.shape
        ds, dp, do = torch.empty_like(s), torch.empty_like(p), torch.empty_like(o)
        g = gout.contiguous().float()
        lib.call("kgeb_score_spo_bwd", ctx.model_id, ctx.l_norm, s.data_ptr(), None, p.data_ptr(), None,
                 o.data_ptr(), None, 0, n, d, lib.f32(g, "grad"), ds.data_ptr(), dp.data_ptr(), do.data_ptr(),
                 lib.stream_ptr(s))
        return None, None, ds, dp, do


def score_spo_emb(model: str, l_norm, s_emb, p_emb, o_emb) -> torch.Tensor:
    d = s_emb.shape[1]
    if p_emb.shape[1] != relation_dim(model, d):
        raise ValueError(f"{model}: relation embeddings must have {relation_dim(model, d)} columns "
                         f"(got {p_emb.shape[1]})")
    return _ScoreSpoEmb.apply(lib.MODELS[model], int(l_norm), s_emb, p_emb, o_emb)


class _ScoreSpoIndexed(torch.autograd.Function):
    """KgeModel.score_spo with the three gathers fused into the scorer (no [n,d] intermediates forward)."""

    @staticmethod
    def forward(ctx, model_id, l_norm, ent_w, rel_w, s, p, o):
        ew, rw = ent_w.detach(), rel_w.detach()
        s, p, o = lib.same_index_width(s.reshape(-1), p.reshape(-1), o.reshape(-1))
        n, d = s.numel(), ew.shape[1]
        out = torch.empty(n, dtype=torch.float32, device=ew.device)
        lib.call("kgeb_score_spo", model_id, l_norm, lib.f32(ew, "entity table"), s.data_ptr(),
                 lib.f32(rw, "relation table"), p.data_ptr(), ew.data_ptr(), o.data_ptr(),
                 int(s.dtype == torch.int64), n, d, out.data_ptr(), lib.stream_ptr(ew))
        ctx.save_for_backward(ew, rw, s, p, o)
        ctx.model_id, ctx.l_norm = model_id, l_norm
        return out

    @staticmethod
    def backward(ctx, gout):
        ew, rw, s, p, o = ctx.saved_tensors
        n, d, dr = s.numel(), ew.shape[1], rw.shape[1]
        dev = ew.device
        dso = torch.empty(2 * n, d, dtype=torch.float32, device=dev)  # rows of s then rows of o
        dp = torch.empty(n, dr, dtype=torch.float32, device=dev)
        g = gout.contiguous().float()
        lib.call("kgeb_score_spo_bwd", ctx.model_id, ctx.l_norm, ew.data_ptr(), s.data_ptr(), rw.data_ptr(),
                 p.data_ptr(), ew.data_ptr(), o.data_ptr(), int(s.dtype == torch.int64), n, d, lib.f32(g, "grad"),
                 dso.data_ptr(), dp.data_ptr(), dso[n:].data_ptr(), lib.stream_ptr(ew))
        g_ent = torch.zeros_like(ew)
        scatter_add_rows_(g_ent, torch.cat((s, o)), dso)
        g_rel = torch.zeros_like(rw)
        scatter_add_rows_(g_rel, p, dp)
        return None, None, g_ent, g_rel, None, None, None


def score_spo_indexed(model: str, l_norm, ent_w, rel_w, s, p, o) -> torch.Tensor:
    return _ScoreSpoIndexed.apply(lib.MODELS[model], int(l_norm), ent_w, rel_w, s, p, o)


# ---------------------------------------------------------------------------------------------
# query transforms of the sp_/_po forms
# ---------------------------------------------------------------------------------------------
class _QueryBuild(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model_id, combine, a_emb, p_emb):
        a, p = a_emb.detach().contiguous(), p_emb.detach().contiguous()
        n, d = a.shape
        q = torch.empty(n, d, dtype=torch.float32, device=a.device)
        lib.call("kgeb_query_build", model_id, combine, None, lib.f32(a, "embeddings"), None, lib.f32(p, "p_emb"), None, 0,
                 n, d, q.data_ptr(), lib.stream_ptr(a))
        ctx.save_for_backward(a, p)
        ctx.model_id, ctx.combine = model_id, combine
        return q

    @staticmethod
    def backward(ctx, dq):
        a, p = ctx.saved_tensors
        n, d = a.shape
        da, dp = torch.empty_like(a), torch.empty_like(p)
        g = dq.contiguous().float()
        lib.call("kgeb_query_bwd", ctx.model_id, ctx.combine, None, a.data_ptr(), None, p.data_ptr(), None, 0, n, d,
                 lib.f32(g, "grad"), da.data_ptr(), dp.data_ptr(), lib.stream_ptr(a))
        return None, None, da, dp


def query_build(model: str, combine: int, a_emb, p_emb) -> torch.Tensor:
    d = a_emb.shape[1]
    if p_emb.shape[1] != relation_dim(model, d):
        raise ValueError(f"{model}: relation embeddings must have {relation_dim(model, d)} columns "
                         f"(got {p_emb.shape[1]})")
    if a_emb.shape[0] != p_emb.shape[0]:
        raise ValueError("query embeddings must have the same number of rows")
    return _QueryBuild.apply(lib.MODELS[model], combine, a_emb, p_emb)


# ---------------------------------------------------------------------------------------------
# K4 / K6 / K7: materialised all-candidate scores
# ---------------------------------------------------------------------------------------------
class _ScoreAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, math, q, cand):
        qd, cd = q.detach().contiguous(), cand.detach().contiguous()
        b, d = qd.shape
        m = cd.shape[0]
        out = torch.empty(b, m, dtype=torch.float32, device=qd.device)
        lib.call("kgeb_score_all", kind, math, lib.f32(qd, "queries"), b, d, lib.f32(cd, "candidates"), None, 0, m,
                 out.data_ptr(), m, 0, lib.stream_ptr(qd))
        ctx.save_for_backward(qd, cd, out if kind in (lib.NEG_L2, lib.ROT_L2) else None)
        ctx.kind = kind
        return out

    @staticmethod
    def backward(ctx, g):
        qd, cd, x = ctx.saved_tensors
        b, d = qd.shape
        m = cd.shape[0]
        g = g.contiguous().float()
        dq = torch.empty_like(qd) if ctx.needs_input_grad[2] else None
        dc = torch.empty_like(cd) if ctx.needs_input_grad[3] else None
        lib.call("kgeb_score_all_bwd", ctx.kind, qd.data_ptr(), b, d, cd.data_ptr(), None, 0, m, lib.f32(g, "grad"),
                 None if x is None else x.data_ptr(), m, 0, None if dq is None else dq.data_ptr(),
                 None if dc is None else dc.data_ptr(), lib.stream_ptr(qd))
        return None, None, dq, dc


def score_all(kind: int, q: torch.Tensor, cand: torch.Tensor, math: int = lib.MATH_FP32) -> torch.Tensor:
    if q.shape[1] != cand.shape[1]:
        raise ValueError("query and candidate embeddings must have the same width")
    return _ScoreAll.apply(kind, math, q, cand)


def score_all_into(kind: int, q: torch.Tensor, table: torch.Tensor, out: torch.Tensor, col_off: int,
                   cand_idx: Optional[torch.Tensor] = None, math: int = lib.MATH_FP32) -> None:
    """No-grad form writing into columns [col_off, col_off+m) of a wider matrix (score_sp_po, kge_model.py:706-746)."""
    b, d = q.shape
    ip, i64 = lib.idx(cand_idx)
    m = table.shape[0] if cand_idx is None else cand_idx.numel()
    lib.call("kgeb_score_all", kind, math, lib.f32(q, "queries"), b, d, lib.f32(table, "table"), ip, i64, m,
             lib.f32(out, "out"), out.shape[1], col_off, lib.stream_ptr(q))


# ---------------------------------------------------------------------------------------------
# negative-sampling pairs (train.py:872-893 without the expansion to B*(1+N) triples)
# ---------------------------------------------------------------------------------------------
class _PairsScore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, q, table, cand, sparse):
        qd, td = q.detach().contiguous(), table.detach()
        cand = cand.contiguous()
        b, m = cand.shape
        d = qd.shape[1]
        out = torch.empty(b, m, dtype=torch.float32, device=qd.device)
        ip, i64 = lib.idx(cand)
        lib.call("kgeb_pairs_score", kind, lib.f32(qd, "queries"), lib.f32(td, "table"), ip, i64, b, m, d,
                 out.data_ptr(), lib.stream_ptr(qd))
        ctx.save_for_backward(qd, td, cand, out)
        ctx.kind, ctx.sparse = kind, sparse
        return out

    @staticmethod
    def backward(ctx, g):
        qd, td, cand, out = ctx.saved_tensors
        b, m = cand.shape
        d = qd.shape[1]
        g = g.contiguous().float()
        dq = torch.empty_like(qd)
        dc = torch.empty(b * m, d, dtype=torch.float32, device=qd.device)
        ip, i64 = lib.idx(cand)
        lib.call("kgeb_pairs_bwd", ctx.kind, qd.data_ptr(), td.data_ptr(), ip, i64, b, m, d, lib.f32(g, "grad"),
                 out.data_ptr(), dq.data_ptr(), dc.data_ptr(), lib.stream_ptr(qd))
        if ctx.sparse:
            ids, rows, cnt = segment_reduce_rows(cand.reshape(-1), dc)
            k = int(cnt.item())
            gt = torch.sparse_coo_tensor(ids[:k].unsqueeze(0), rows[:k], tuple(td.shape), device=td.device)
        else:
            gt = torch.zeros_like(td)
            scatter_add_rows_(gt, cand.reshape(-1), dc)
        return None, dq, gt, None, None


def pairs_score(kind: int, q: torch.Tensor, table: torch.Tensor, cand: torch.Tensor, sparse: bool = False):
    return _PairsScore.apply(kind, q, table, cand, sparse)


# ---------------------------------------------------------------------------------------------
# Lp penalties (lookup_embedder.py:112-158) -- csrc/penalty.cu
# ---------------------------------------------------------------------------------------------
def _penalty_ws(device, n_indexes: int, numel: int) -> torch.Tensor:
    return _workspace(device, lib.load().kgeb_penalty_workspace_bytes(int(n_indexes), int(numel)))


class _LpPenaltyDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, p, reg_weight):
        w = weight.detach()
        value = torch.empty(1, dtype=torch.float32, device=w.device)
        ws = _penalty_ws(w.device, 0, w.numel())
        lib.call("kgeb_lp_penalty_dense", lib.f32(w, "weight"), w.numel(), int(p), float(reg_weight), None,
                 value.data_ptr(), ws.data_ptr(), ws.numel(), lib.stream_ptr(w))
        ctx.save_for_backward(weight)
        ctx.p, ctx.reg_weight = int(p), float(reg_weight)
        return value.view(())

    @staticmethod
    def backward(ctx, gout):
        (weight,) = ctx.saved_tensors
        w = weight.detach()
        grad = torch.zeros_like(w)
        value = torch.empty(1, dtype=torch.float32, device=w.device)
        ws = _penalty_ws(w.device, 0, w.numel())
        lib.call("kgeb_lp_penalty_dense", lib.f32(w, "weight"), w.numel(), ctx.p, ctx.reg_weight, grad.data_ptr(),
                 value.data_ptr(), ws.data_ptr(), ws.numel(), lib.stream_ptr(w))
        return grad * gout, None, None


class _LpPenaltyRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, indexes, p, reg_weight):
        w = weight.detach()
        indexes = indexes.contiguous()          # (a column view of the triples: keep the copy alive across the call)
        iptr, i64 = lib.idx(indexes, "indexes")
        value = torch.empty(1, dtype=torch.float32, device=w.device)
        ws = _penalty_ws(w.device, indexes.numel(), 0)
        lib.call("kgeb_lp_penalty_rows", lib.f32(w, "weight"), w.shape[0], w.shape[1], iptr, i64, indexes.numel(), int(p),
                 float(reg_weight), None, value.data_ptr(), ws.data_ptr(), ws.numel(), lib.stream_ptr(w))
        ctx.save_for_backward(weight, indexes)
        ctx.p, ctx.reg_weight = int(p), float(reg_weight)
        return value.view(())

    @staticmethod
    def backward(ctx, gout):
        weight, indexes = ctx.saved_tensors
        w = weight.detach()
        iptr, i64 = lib.idx(indexes, "indexes")     # made contiguous in forward
        grad = torch.zeros_like(w)
        value = torch.empty(1, dtype=torch.float32, device=w.device)
        ws = _penalty_ws(w.device, indexes.numel(), 0)
        lib.call("kgeb_lp_penalty_rows", lib.f32(w, "weight"), w.shape[0], w.shape[1], iptr, i64, indexes.numel(), ctx.p,
                 ctx.reg_weight, grad.data_ptr(), value.data_ptr(), ws.data_ptr(), ws.numel(), lib.stream_ptr(w))
        return grad * gout, None, None, None


def lp_penalty(weight: torch.Tensor, p: int, reg_weight: float, indexes: Optional[torch.Tensor] = None) -> torch.Tensor:
    """reg_weight/p * ||W||_p^p (indexes None: unweighted, whole table) or the weighted per-distinct-index form divided
    by len(indexes) (lookup_embedder.py:126-155); differentiable w.r.t. `weight`."""
    if not weight.is_cuda:
        raise ValueError("weight must be a CUDA tensor (no CPU path exists)")
    if int(p) != p or p < 1:
        raise NotImplementedError(f"regularize_args.p={p}: integer p >= 1 only")
    if indexes is None:
        return _LpPenaltyDense.apply(weight, int(p), float(reg_weight))
    if indexes.numel() == 0:
        return weight.new_zeros(())
    return _LpPenaltyRows.apply(weight, indexes, int(p), float(reg_weight))
