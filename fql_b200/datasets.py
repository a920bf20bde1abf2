"""Device-resident mirror of the reference sampler (utils/datasets.py): Dataset / ReplayBuffer with the same
attributes (`size`, `frame_stack`, `p_aug`, `terminal_locs`, `initial_locs`) and the same RNG draw order on the GLOBAL
numpy MT19937 (randint(size, B) -> [rand() -> randint(0, 7, (B,2))]), so sampled indices and crops are bit-identical
to the reference given the same np.random.seed.  The arrays live in HBM; the fancy-index gather, frame stacking and
edge-padded crop run as CUDA kernels (fql_gather_rows / fql_gather_frames) and the batch never visits the host.
No CPU path: the gather is always the CUDA kernel.

`sample` leaves the step: the index upload (a ring of pinned staging buffers, never waited on until it wraps) and the gather
kernels are enqueued on the dataset's own stream, so the batch of update k+1 is assembled while update k is still running; the
caller's stream only waits (on the device) for the event behind the gathers.  The host never blocks and the RNG draw order is
untouched (nothing is drawn ahead of the reference's call order).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Dataset:
    """utils/datasets.py:36-112 (flat dict of arrays)."""

    def __init__(self, data, device=None, size=None):
        assert 'observations' in data
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self._dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in data.items()}
        self.size = max(len(v) for v in data.values()) if size is None else size   # get_size, datasets.py:11-14
        self.frame_stack = None
        self.p_aug = None
        self.return_next_actions = False
        term = np.asarray(data['terminals'])
        self.terminal_locs = np.nonzero(term > 0)[0]                               # datasets.py:61
        self.initial_locs = np.concatenate([[0], self.terminal_locs[:-1] + 1])     # datasets.py:62
        self._lib = _lib.lib()
        self._pins, self._pin_i = [], 0
        self._side = None            # the sampler's own stream (created on first use)

    @classmethod
    def create(cls, freeze=True, device=None, **fields):
        return cls(fields, device=device)

    def __getitem__(self, k):
        return self._dev[k]

    def keys(self):
        return self._dev.keys()

    def get_random_idxs(self, num_idxs):
        return np.random.randint(self.size, size=num_idxs)                          # datasets.py:66

    # -- device helpers ------------------------------------------------------------------------------------
    _PIN_RING = 16

    def _to_dev(self, arr):
        """int64 indices -> device through the next slot of a ring of pinned staging buffers: a slot is only waited on when the
        ring wraps around to a copy that has not run yet (the host may be many steps ahead of the GPU)."""
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64))
        n = t.numel()
        i = self._pin_i % self._PIN_RING
        self._pin_i += 1
        if i >= len(self._pins):
            self._pins.append([torch.empty(max(n, 4096), dtype=torch.int64).pin_memory(), torch.cuda.Event(), False])
        slot = self._pins[i]
        if slot[2]:
            slot[1].synchronize()
        if slot[0].numel() < n:
            slot[0] = torch.empty(n, dtype=torch.int64).pin_memory()
        slot[0][:n].copy_(t.reshape(-1))
        d = torch.empty(t.shape, dtype=torch.int64, device=self.device)
        d.reshape(-1).copy_(slot[0][:n], non_blocking=True)
        slot[1].record(torch.cuda.current_stream(self.device))
        slot[2] = True
        return d

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _gather(self, arr, idx_dev):
        n = idx_dev.numel()
        out = torch.empty((n,) + tuple(arr.shape[1:]), dtype=arr.dtype, device=self.device)
        if n:
            row_bytes = arr[0].numel() * arr.element_size() if arr.dim() > 1 else arr.element_size()
            _lib.check(self._lib.fql_gather_rows(_p(arr), _p(out), _p(idx_dev), n, row_bytes, self._stream()), 'fql_gather_rows')
        return out

    # -- Dataset API ---------------------------------------------------------------------------------------
    def get_subset(self, idxs, idx_dev=None, skip=()):
        idx_dev = self._to_dev(idxs) if idx_dev is None else idx_dev
        result = {k: self._gather(v, idx_dev) for k, v in self._dev.items() if k not in skip}
        if self.return_next_actions:
            nxt = self._to_dev(np.minimum(np.asarray(idxs) + 1, self.size - 1))
            result['next_actions'] = self._gather(self._dev['actions'], nxt)
        return result

    def sample(self, batch_size, idxs=None):
        """datasets.py:68-92.  Returns a dict of DEVICE tensors with the reference's dtypes and shapes."""
        if idxs is None:
            idxs = self.get_random_idxs(batch_size)
        idxs = np.asarray(idxs)
        caller = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._done = torch.cuda.Event()
        with torch.cuda.device(self.device), torch.cuda.stream(self._side):
            idx_dev = self._to_dev(idxs)
            obs = self._dev['observations']
            is_img = obs.dim() == 4
            init_dev = None
            if self.frame_stack is not None:
                initial_state_idxs = self.initial_locs[np.searchsorted(self.initial_locs, idxs, side='right') - 1]  # :75
                init_dev = self._to_dev(initial_state_idxs)
            crop = None
            if self.p_aug is not None:
                if np.random.rand() < self.p_aug:                                                    # datasets.py:90
                    crop = np.random.randint(0, 2 * 3 + 1, (len(idxs), 2))                            # datasets.py:106
            fused = is_img and (self.frame_stack is not None or crop is not None)
            batch = self.get_subset(idxs, idx_dev, skip=('observations', 'next_observations') if fused else ())
            if fused:
                batch['observations'], batch['next_observations'] = self._frames(idx_dev, init_dev, crop, len(idxs))
            elif self.frame_stack is not None:
                batch['observations'], batch['next_observations'] = self._stack_state(idxs, initial_state_idxs, idx_dev)
            self._done.record(self._side)
        caller.wait_event(self._done)          # device-side: the consumer's stream runs behind the gathers, the host does not wait
        for v in batch.values():
            v.record_stream(caller)            # the caching allocator must not recycle a batch the caller's stream still reads
        return batch

    def augment(self, batch, keys):
        raise NotImplementedError('augmentation is fused into sample(): set p_aug (datasets.py:88-91)')

    def _frames(self, idx_dev, init_dev, crop, n):
        obs, nobs = self._dev['observations'], self._dev['next_observations']
        _, H, W, Cc = obs.shape
        fs = self.frame_stack if self.frame_stack is not None else 1
        o = torch.empty((n, H, W, fs * Cc), dtype=torch.uint8, device=self.device)
        no = torch.empty_like(o)
        crop_dev = self._to_dev(crop) if crop is not None else None
        if n:
            _lib.check(self._lib.fql_gather_frames(_p(obs), _p(nobs), _p(o), _p(no), _p(idx_dev), _p(init_dev), _p(crop_dev), n,
                                                   H, W, Cc, fs, 3, self._stream()), 'fql_gather_frames')
        return o, no

    def _stack_state(self, idxs, init, idx_dev):
        """frame_stack on vector observations: concat along the last axis (datasets.py:78-87)."""
        fs = self.frame_stack
        obs, nobs = [], []
        for i in reversed(range(fs)):
            cur = self._to_dev(np.maximum(idxs - i, init))
            g = self._gather(self._dev['observations'], cur)
            obs.append(g)
            if i != fs - 1:
                nobs.append(g)
        nobs.append(self._gather(self._dev['next_observations'], idx_dev))
        return torch.cat(obs, -1), torch.cat(nobs, -1)


class ReplayBuffer(Dataset):
    """utils/datasets.py:435-495 (offline use: create_from_initial_dataset; add_transition is online RL, out of scope)."""

    @classmethod
    def create_from_initial_dataset(cls, init_dataset, size, device=None):
        n = max(len(v) for v in init_dataset.values())
        buf = {}
        for k, v in init_dataset.items():
            v = np.asarray(v)
            b = np.zeros((size, *v.shape[1:]), dtype=v.dtype)
            b[: len(v)] = v
            buf[k] = b
        ds = cls(buf, device=device)
        ds.max_size = size
        ds.size = ds.pointer = n
        return ds

    def add_transition(self, transition):
        raise NotImplementedError('online replay insertion is outside the offline update hot path (SURVEY 2)')


def _dataset_from_initial(cls, init_dataset, size, device=None):
    return ReplayBuffer.create_from_initial_dataset(init_dataset, size, device)


Dataset.create_from_initial_dataset = classmethod(_dataset_from_initial)
