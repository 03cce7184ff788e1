"""KvsAllIndex as CSR arrays (kge/indexing.py:8-139), built on the host like the reference and
mirrored on the device so that label / filter coordinates are looked up by a kernel
(kgeb_csr_lookup) instead of the Python loops of train.py:590-677 and job/util.py:5-38."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib

_KEYS = {"sp": ((0, 1), 2), "po": ((1, 2), 0), "so": ((0, 2), 1)}


class KvsAllIndex:
    """keys [K,2] (lexicographic), offsets [K+1], values [N] (ascending within a key; duplicates kept)."""

    def __init__(self, triples, key: str):
        key_cols, value_col = _KEYS[key]
        t = np.asarray(triples.cpu() if isinstance(triples, torch.Tensor) else triples)
        # stable multi-key sort = lexsort by (value, key1, key0); same order as indexing.py:86-98
        order = np.lexsort((t[:, value_col], t[:, key_cols[1]], t[:, key_cols[0]]))
        t = t[order]
        # unique keys + first row of each: the rows are sorted, so a key starts wherever it differs from the row above
        # (what np.unique(axis=0, return_index=True) returns, indexing.py:41-44, without its void-dtype sort: 20 M triples
        # of the Wikidata5M shape take seconds instead of minutes)
        kk = t[:, list(key_cols)]
        if len(kk):
            start = np.ones(len(kk), dtype=bool)
            start[1:] = (kk[1:] != kk[:-1]).any(axis=1)
            first = np.flatnonzero(start)
            keys = kk[first]
        else:
            first, keys = np.zeros(0, dtype=np.int64), kk.reshape(0, 2)
        self.key = key
        self._keys = torch.from_numpy(np.ascontiguousarray(keys)).long()
        self._values_offset = torch.from_numpy(np.append(first, len(t)).astype(np.int64))
        self._values = torch.from_numpy(np.ascontiguousarray(t[:, value_col])).long()
        self._dev: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self._index_of_key: Optional[dict] = None

    @classmethod
    def from_arrays(cls, keys, values_offset, values, key: str = "sp") -> "KvsAllIndex":
        """Adopts the arrays of an index that already exists (the reference's own KvsAllIndex keeps the same three
        arrays, indexing.py:36-55: `_keys` lexicographic, `_values_offset`, `_values` ascending within a key) instead of
        sorting the triples again."""
        self = cls.__new__(cls)
        self.key = key
        self._keys = torch.as_tensor(keys).long().contiguous().view(-1, 2)
        self._values_offset = torch.as_tensor(values_offset).long().contiguous()
        self._values = torch.as_tensor(values).long().contiguous()
        self._dev = {}
        self._index_of_key = None
        return self

    def __len__(self):
        return len(self._keys)

    # dict-style access as in the reference (indexing.py:57-84) -- host side, for small uses/tests
    def get(self, key, default=None):
        if self._index_of_key is None:
            self._index_of_key = {(int(a), int(b)): i for i, (a, b) in enumerate(self._keys.tolist())}
        i = self._index_of_key.get((int(key[0]), int(key[1])))
        if i is None:
            return torch.zeros(0, dtype=torch.long) if default is None else default
        return self._values[self._values_offset[i]:self._values_offset[i + 1]]

    __getitem__ = get

    def device_arrays(self, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        k = str(device)
        if k not in self._dev:
            self._dev[k] = (self._keys.to(device).contiguous(), self._values_offset.to(device),
                            self._values.to(device))
        return self._dev[k]

    def lookup(self, pairs: torch.Tensor) -> torch.Tensor:
        """Row of each (k1,k2) pair in the index or -1; `pairs` is a [n,2] int64 CUDA tensor."""
        keys, _, _ = self.device_arrays(pairs.device)
        pairs = pairs.long().contiguous()
        out = torch.empty(pairs.shape[0], dtype=torch.int64, device=pairs.device)
        lib.call("kgeb_csr_lookup", keys.data_ptr(), keys.shape[0], pairs.data_ptr(), pairs.shape[0], out.data_ptr(),
                 lib.stream_ptr(pairs))
        return out


def filter_csr(sp_indexes: Sequence["KvsAllIndex"], po_indexes: Sequence["KvsAllIndex"], s, p, o):
    """Filter CSR of an evaluation batch, built on the device (kgeb_filter_csr_count / _fill): rows 0..B-1 = known
    objects of (s,p), rows B..2B-1 = known subjects of (p,o), over all given splits, ascending per row (duplicates
    kept; the ranking kernels skip repeated columns).  Replaces EntityRankingJob._collate's dict lookups
    (entity_ranking.py:53-77, job/util.py:5-38).  One host sync (the total size of the result)."""
    dev = s.device
    k, b = len(sp_indexes), s.numel()
    sp = lib.index_descs([ix.device_arrays(dev) for ix in sp_indexes])
    po = lib.index_descs([ix.device_arrays(dev) for ix in po_indexes])
    sp_ptr, s64 = lib.idx(s.contiguous())
    if p.dtype != s.dtype or o.dtype != s.dtype:
        raise ValueError("s, p, o must have the same index dtype")
    st = lib.stream_ptr(s)
    n = 2 * b * k
    src_row = torch.empty(n, dtype=torch.int64, device=dev)
    pos = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    lib.call("kgeb_filter_csr_count", sp, po, k, sp_ptr, p.contiguous().data_ptr(), o.contiguous().data_ptr(), s64, b,
             src_row.data_ptr(), pos[1:].data_ptr(), st)
    torch.cumsum(pos[1:], 0, out=pos[1:])
    off = pos[::k].contiguous()
    total = int(pos[-1].item())
    col = torch.empty(total, dtype=torch.int64, device=dev)
    if total == 0:
        return off, col
    if k == 1:     # a single sorted list per row
        lib.call("kgeb_filter_csr_fill", sp, po, k, b, src_row.data_ptr(), pos.data_ptr(), col.data_ptr(), None, st)
        return off, col
    lib.call("kgeb_filter_csr_fill", sp, po, k, b, src_row.data_ptr(), pos.data_ptr(), None, col.data_ptr(), st)
    col = torch.sort(col).values          # keys (row << 32 | entity): merges the per-split lists of every row
    return off, col.bitwise_and_(0xFFFFFFFF)


def kvsall_batch(sp_index: "KvsAllIndex", po_index: "KvsAllIndex", example_ids: torch.Tensor, out=None, capacity=None):
    """KvsAll training batch on the device from example ids (kgeb_kvsall_batch_count / _fill; train.py:590-677):
    returns (a_idx, p_idx, row_combine, lab_off, lab_col, overflow).  With `out` = (a_idx, p_idx, row_combine, lab_off,
    lab_col) preallocated static buffers nothing is allocated and nothing syncs; lab_col then has `capacity` entries and
    the int32 device flag `overflow` reports a batch that did not fit."""
    dev = example_ids.device
    b = example_ids.numel()
    sp = lib.index_descs([sp_index.device_arrays(dev)])
    po = lib.index_descs([po_index.device_arrays(dev)])
    st = lib.stream_ptr(example_ids)
    ids = lib.i64(example_ids.contiguous(), "example ids")
    if out is None:
        a = torch.empty(b, dtype=torch.int64, device=dev)
        p = torch.empty(b, dtype=torch.int64, device=dev)
        rc = torch.empty(b, dtype=torch.int32, device=dev)
        lab_off = torch.zeros(b + 1, dtype=torch.int64, device=dev)
    else:
        a, p, rc, lab_off, lab_col = out
        lab_off[:1].zero_()
    lib.call("kgeb_kvsall_batch_count", sp, po, ids, b, a.data_ptr(), p.data_ptr(), rc.data_ptr(), lab_off[1:].data_ptr(), st)
    torch.cumsum(lab_off[1:], 0, out=lab_off[1:])
    if out is None:
        capacity = int(lab_off[-1].item())
        lab_col = torch.empty(capacity, dtype=torch.int64, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.call("kgeb_kvsall_batch_fill", sp, po, ids, b, lab_off.data_ptr(), int(capacity), lab_col.data_ptr(),
             overflow.data_ptr(), st)
    return a, p, rc, lab_off, lab_col, overflow


def gather_csr_rows(offsets: torch.Tensor, values: torch.Tensor, rows: torch.Tensor, add: int = 0):
    """CSR slice for the given index rows (-1 = empty): returns (row_offsets [n+1], concatenated values + add)."""
    valid = rows >= 0
    safe = torch.where(valid, rows, torch.zeros_like(rows))
    start = offsets[safe]
    length = torch.where(valid, offsets[safe + 1] - start, torch.zeros_like(start))
    out_off = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rows.device)
    out_off[1:] = torch.cumsum(length, 0)
    total = int(out_off[-1].item())
    if total == 0:
        return out_off, torch.zeros(0, dtype=torch.int64, device=rows.device)
    seg = torch.repeat_interleave(torch.arange(rows.numel(), device=rows.device), length, output_size=total)
    pos = torch.arange(total, device=rows.device) - out_off[seg] + start[seg]
    return out_off, values[pos] + add


def merge_sorted_csr(parts: Sequence[Tuple[torch.Tensor, torch.Tensor]], num_rows: int):
    """Union of several per-row CSR lists -> one CSR with ascending columns per row (duplicates kept; the
    ranking kernels skip repeated columns)."""
    parts = [p for p in parts if p is not None]
    dev = parts[0][0].device
    rows, cols = [], []
    for off, col in parts:
        n = off[1:] - off[:-1]
        rows.append(torch.repeat_interleave(torch.arange(num_rows, device=dev), n, output_size=col.numel()))
        cols.append(col)
    r, c = torch.cat(rows), torch.cat(cols)
    order = torch.argsort((r << 32) + c)
    r, c = r[order], c[order]
    off = torch.zeros(num_rows + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(torch.bincount(r, minlength=num_rows), 0)
    return off, c.contiguous()


# ---------------------------------------------------------------------------------------------
# drill-down indexes of the evaluation job (kge/indexing.py:142-263); built once on the host like the reference does
# ---------------------------------------------------------------------------------------------
def relation_types(train_triples, num_relations: int) -> List[str]:
    """"1-1" / "1-N" / "M-1" / "M-N" per relation (indexing.py:142-179, after Bordes et al. 2013): M if a (p, o) pair of
    the relation has more than 1.5 subjects on average, N if an (s, p) pair has more than 1.5 objects on average."""
    t = np.asarray(train_triples.cpu() if isinstance(train_triples, torch.Tensor) else train_triples).astype(np.int64)
    out = []
    po = np.unique(t[:, [1, 2]], axis=0)
    sp = np.unique(t[:, [0, 1]], axis=0)
    n_triples = np.bincount(t[:, 1], minlength=num_relations).astype(np.float32)
    n_po = np.bincount(po[:, 0], minlength=num_relations).astype(np.float32)
    n_sp = np.bincount(sp[:, 1], minlength=num_relations).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        is_m = (n_triples / n_po) > 1.5        # 0/0 = nan -> False, as in the reference's float tensor
        is_n = (n_triples / n_sp) > 1.5
    for i in range(num_relations):
        out.append("{}-{}".format("M" if is_m[i] else "1", "N" if is_n[i] else "1"))
    return out


def relations_per_type(train_triples, num_relations: int) -> Dict[str, List[int]]:
    """indexing.py:182-197."""
    res: Dict[str, List[int]] = {}
    for i, k in enumerate(relation_types(train_triples, num_relations)):
        res.setdefault(k, []).append(i)
    return res


def frequency_percentiles(train_triples, num_entities: int, num_relations: int) -> Dict[str, Dict[str, List[int]]]:
    """indexing.py:200-263: ids sorted by their training frequency in the slot (ascending, ties by id), cut at 25 / 50 /
    75 % of the vocabulary: {"subject" | "relation" | "object": {"25%" | "50%" | "75%" | "top": ids}}."""
    t = np.asarray(train_triples.cpu() if isinstance(train_triples, torch.Tensor) else train_triples).astype(np.int64)
    res: Dict[str, Dict[str, List[int]]] = {}
    for arg, col, num in (("subject", 0, num_entities), ("relation", 1, num_relations), ("object", 2, num_entities)):
        order = np.argsort(np.bincount(t[:, col], minlength=num), kind="stable")
        res[arg] = {}
        for perc, (begin, end) in (("25%", (0.0, 0.25)), ("50%", (0.25, 0.5)), ("75%", (0.5, 0.75)), ("top", (0.75, 1.0))):
            res[arg][perc] = sorted(int(x) for x in order[int(begin * num):int(end * num)])
    return res
