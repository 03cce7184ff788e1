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
        keys, first = np.unique(t[:, list(key_cols)], axis=0, return_index=True)
        self.key = key
        self._keys = torch.from_numpy(np.ascontiguousarray(keys)).long()
        self._values_offset = torch.from_numpy(np.append(first, len(t)).astype(np.int64))
        self._values = torch.from_numpy(np.ascontiguousarray(t[:, value_col])).long()
        self._dev: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self._index_of_key: Optional[dict] = None

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
