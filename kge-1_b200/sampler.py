"""Negative samplers on the device (kge/util/sampler.py): same options, same `sample(positive_triples, slot,
num_samples)` call and result layout (`[batch, num_samples]` int64), drawn by the Philox kernels of
csrc/sampler.cu instead of torch.randint / numpy / numba on the CPU inside DataLoader workers.

The random stream necessarily differs from the reference's; what is kept is the contract (sampler.py:78-120):
uniform over the slot's vocabulary, optional filtering of known positives with redraws, optional shared sampling
(without replacement, or with replacement via the reference's distinct-count construction)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import lib
from .index import KvsAllIndex

S, P, O = 0, 1, 2
SLOT_STR = ["s", "p", "o"]
_PAIR = ["po", "so", "sp"]                   # key pair of the filter index per slot (sampler.py:44, 153)
_PAIR_COLS = [[P, O], [S, O], [S, P]]


class KgeSampler:
    """sampler.py:15-146 (configuration, `sample`)."""

    def __init__(self, num_entities: int, num_relations: int, num_samples: Sequence[int] = (-1, 0, -1),
                 filter_positives: Sequence[bool] = (False, False, False), shared: bool = False,
                 with_replacement: bool = True, filter_indexes: Optional[Dict[int, KvsAllIndex]] = None,
                 train_triples=None, seed: int = 0, device="cuda"):
        if not with_replacement and not shared:
            raise ValueError("Without replacement sampling is only supported when shared negative sampling is enabled.")
        if any(filter_positives) and shared:
            raise ValueError("Filtering is not supported when shared negative sampling is enabled.")
        self.shared, self.with_replacement = shared, with_replacement
        self.vocabulary_size = [num_entities, num_relations, num_entities]
        self.filter_positives = list(filter_positives)
        ns = list(num_samples)
        for slot, copy_from in [(S, O), (P, None), (O, S)]:          # auto config, sampler.py:59-65
            if ns[slot] < 0:
                ns[slot] = ns[copy_from] if copy_from is not None and ns[copy_from] > 0 else 0
        self.num_samples = ns
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("the sampler draws on the device (no CPU path exists)")
        self.filter_indexes: Dict[int, KvsAllIndex] = dict(filter_indexes or {})
        for slot in (S, P, O):
            if self.filter_positives[slot] and slot not in self.filter_indexes:
                if train_triples is None:
                    raise ValueError(f"filtering slot {SLOT_STR[slot]} needs its index or the filtering split's triples")
                self.filter_indexes[slot] = KvsAllIndex(train_triples, _PAIR[slot])
        # Philox state {seed, offset} in device memory: a graph replay of sample() draws fresh numbers
        self.state = torch.tensor([seed, 0], dtype=torch.int64, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.meta = torch.zeros(2, dtype=torch.int32, device=self.device)
        self._ws: Optional[torch.Tensor] = None

    @staticmethod
    def create(sampling_type: str, *args, **kw) -> "KgeSampler":
        if sampling_type == "uniform":
            return KgeUniformSampler(*args, **kw)
        raise ValueError("negative_sampling.sampling_type")     # sampler.py:67-77 (frequency: see DESIGN.md)

    def sample(self, positive_triples: torch.Tensor, slot: int, num_samples: Optional[int] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`out`: optional preallocated [batch, num_samples] int64 CUDA buffer (the static input of a captured step)."""
        if num_samples is None:
            num_samples = self.num_samples[slot] if out is None else out.shape[1]
        pt = positive_triples
        if not pt.is_cuda:
            raise ValueError("positive_triples must be a CUDA tensor (no CPU path exists)")
        pt = pt.long().contiguous()
        if out is None:
            out = torch.empty(pt.shape[0], num_samples, dtype=torch.int64, device=pt.device)
        elif (not out.is_cuda or out.dtype != torch.int64 or not out.is_contiguous()
              or tuple(out.shape) != (pt.shape[0], num_samples)):
            raise ValueError("out must be a contiguous [batch, num_samples] int64 CUDA tensor")
        if self.shared and num_samples + 1 > self.vocabulary_size[slot]:
            raise ValueError(f"{num_samples + 1} distinct shared samples from a vocabulary of {self.vocabulary_size[slot]}")
        if out.numel() == 0:                   # empty batch / no samples for this slot: nothing to draw
            return out
        if self.shared:
            neg = self._sample_shared(pt, slot, num_samples, out)
        else:
            neg = self._sample(pt, slot, num_samples, out)
        if self.filter_positives[slot]:
            neg = self._filter_and_resample(neg, slot, pt)
        lib.call("kgeb_philox_advance", self.state.data_ptr(), 1 << 20, lib.stream_ptr(pt))
        return neg

    def check_status(self):
        """Host check (one sync) that no row was unsatisfiable / no shared draw failed to converge."""
        if int(self.status.item()) != 0 or int(self.meta[1].item()) != 0:
            raise RuntimeError("negative sampling could not find enough true negatives")


class KgeUniformSampler(KgeSampler):
    """sampler.py:190-315."""

    def _sample(self, pt, slot, num_samples, out):
        b = pt.shape[0]
        lib.call("kgeb_sample_uniform", self.state.data_ptr(), self.vocabulary_size[slot], b * num_samples, out.data_ptr(),
                 lib.stream_ptr(pt))
        return out

    def _sample_shared(self, pt, slot, num_samples, out):
        b = pt.shape[0]
        L = lib.load()
        need = L.kgeb_sample_shared_workspace_bytes(num_samples)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=pt.device)
        positives = pt[:, slot].contiguous()
        lib.call("kgeb_sample_shared", self.state.data_ptr(), self.vocabulary_size[slot], positives.data_ptr(), b,
                 num_samples, int(self.with_replacement), out.data_ptr(), self.meta.data_ptr(), self._ws.data_ptr(),
                 self._ws.numel(), lib.stream_ptr(pt))
        return out

    def _filter_and_resample(self, neg, slot, pt):
        ix = self.filter_indexes[slot]
        desc = lib.index_descs([ix.device_arrays(pt.device)])
        cols = _PAIR_COLS[slot]
        ka, kb = pt[:, cols[0]].contiguous(), pt[:, cols[1]].contiguous()
        lib.call("kgeb_sample_filter", self.state.data_ptr(), self.vocabulary_size[slot], desc, ka.data_ptr(), kb.data_ptr(),
                 pt.shape[0], neg.shape[1], neg.data_ptr(), self.status.data_ptr(), lib.stream_ptr(pt))
        return neg
