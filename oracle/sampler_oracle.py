"""CPU oracle of the on-device negative sampler (kge-1_b200/csrc/sampler.cu)  --  TEST INFRASTRUCTURE ONLY.

Two things live here:

1. a numpy restatement of the *device generator*: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy
   as 1, 2, 3", SC'11; checked below against the Random123 known-answer vectors) and the draw / redraw /
   shared-sample procedures built on it.  Given (seed, offset) it reproduces the CUDA kernels bit for bit, which is
   how the tests prove the kernels are deterministic and launch-shape independent.
2. the *reference semantics* the samples must satisfy (kge/util/sampler.py), as predicates:
   `_sample` (sampler.py:195-198) uniform in [0, vocab); `_filter_and_resample` (sampler.py:148-176, 257-315) no
   known positive of the row's key pair survives; `_sample_shared` (sampler.py:200-255) one shared set of
   num_distinct+1 distinct values, each row drops its own positive if present else one random position, WR columns
   beyond num_distinct copy earlier columns.
The reference's own random streams (torch.randint / numpy / random.sample) are not reproducible on a device, so
parity for the sampler is distributional + these predicates; the training path itself is compared with externally
supplied negatives (tests/test_gpu_parity.py).

Only tests/ may import this module.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [n,4] uint32, key [n,2] (or [2]) uint32 -> [n,4] uint32."""
    c = np.array(ctr, dtype=np.uint64).reshape(-1, 4).copy()
    k = np.broadcast_to(np.array(key, dtype=np.uint64).reshape(-1, 2), (c.shape[0], 2)).copy()
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = np.stack([hi1 ^ c[:, 1] ^ k[:, 0], lo1, hi0 ^ c[:, 3] ^ k[:, 1], lo0], 1)
        k = np.stack([(k[:, 0] + np.uint64(W0)) & MASK, (k[:, 1] + np.uint64(W1)) & MASK], 1)
    return c.astype(np.uint32)


def draw(seed: int, offset: int, stream_id: int, elem, t, rng: int) -> np.ndarray:
    """Value(s) in [0, rng): word 0 of Philox at counter (offset + t, elem ^ stream_id << 60), multiply-shift."""
    elem = np.atleast_1d(np.asarray(elem, dtype=np.uint64))
    t = np.broadcast_to(np.asarray(t, dtype=np.uint64), elem.shape)
    ctr64 = (np.uint64(offset) + t)
    ctr = np.stack([ctr64 & MASK, ctr64 >> np.uint64(32), elem & MASK,
                    ((elem >> np.uint64(32)) ^ np.uint64((stream_id << 28) & 0xFFFFFFFF)) & MASK], 1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    w = philox4x32_10(ctr, key)[:, 0].astype(np.uint64)
    return ((w * np.uint64(rng)) >> np.uint64(32)).astype(np.int64)


def sample_uniform(seed: int, offset: int, vocab: int, n: int) -> np.ndarray:
    return draw(seed, offset, 0, np.arange(n), 0, vocab)


def sample_filter(seed: int, offset: int, vocab: int, index: Dict[Tuple[int, int], np.ndarray], key_a, key_b,
                  negatives: np.ndarray, max_tries: int = 1 << 16) -> np.ndarray:
    """index: (a, b) -> ascending array of known positives."""
    out = negatives.copy()
    b, n = out.shape
    for i in range(b):
        pos = index.get((int(key_a[i]), int(key_b[i])))
        if pos is None or len(pos) == 0:
            continue
        pos = set(int(x) for x in pos)
        for j in range(n):
            if int(out[i, j]) not in pos:
                continue
            for t in range(1, max_tries):
                v = int(draw(seed, offset, 1, i * n + j, t, vocab)[0])
                if v not in pos:
                    out[i, j] = v
                    break
    return out


def sample_shared(seed: int, offset: int, vocab: int, positives: np.ndarray, n: int, with_replacement: bool):
    """Returns (samples [B,n], num_distinct, shared [num_distinct+1])."""
    if with_replacement:
        nd = len(np.unique(draw(seed, offset, 2, np.arange(n), 0, vocab - 1)))
    else:
        nd = n
    m = nd + 1
    shared = np.full(m, -1, dtype=np.int64)
    tries = np.zeros(m, dtype=np.int64)
    taken: Dict[int, int] = {}
    pending = list(range(m))
    while pending:
        vals = draw(seed, offset, 3, np.array(pending), tries[pending], vocab)
        winners: Dict[int, int] = {}
        for j, v in zip(pending, vals.tolist()):      # pending is ascending: the smallest position of a round wins
            if v not in taken and v not in winners:
                winners[v] = j
        nxt = []
        for j, v in zip(pending, vals.tolist()):
            if winners.get(v) == j:
                shared[j] = v
            else:
                tries[j] += 1
                nxt.append(j)
        taken.update(winners)
        pending = nxt
    up = np.arange(n)
    if nd < n:
        up[nd:] = draw(seed, offset, 4, np.arange(nd, n), 0, nd)
    pos_of = {int(v): j for j, v in enumerate(shared.tolist())}
    b = len(positives)
    rand_drop = draw(seed, offset, 5, np.arange(b), 0, nd + 1)
    out = np.empty((b, n), dtype=np.int64)
    for i in range(b):
        drop = pos_of.get(int(positives[i]), int(rand_drop[i]))
        row = shared[:nd].copy()
        if drop != nd:
            row[drop] = shared[nd]
        out[i] = row[up]
    return out, nd, shared
