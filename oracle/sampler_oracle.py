"""CPU oracle for the dataset sampler -- TEST INFRASTRUCTURE ONLY (see oracle/fql_oracle.py header).

Literal NumPy transcription of the integer semantics of the reference sampler; the only substitutions are
`jax.tree_util.tree_map` -> dict comprehension (flat dict of arrays) and the jitted `batched_random_crop`
(`jnp.pad(mode='edge')` + `dynamic_slice`) -> `np.pad(mode='edge')` + slice, which is the same integer indexing.

Reference citations (relative to /root/reference):
  utils/datasets.py:53-62    Dataset.__init__   (terminal_locs / initial_locs)
  utils/datasets.py:64-66    get_random_idxs    (global numpy MT19937)
  utils/datasets.py:68-92    sample             (gather, frame stack, p_aug Bernoulli)
  utils/datasets.py:94-100   get_subset
  utils/datasets.py:102-112  augment            (crop_froms = randint(0, 7, (B,2)))
  utils/datasets.py:17-33    random_crop / batched_random_crop
  utils/datasets.py:457-474  ReplayBuffer.create_from_initial_dataset

Draw order on the GLOBAL numpy RNG per sample(B): (1) randint(size, size=B); (2) if p_aug is not None: rand();
(3) only if (2) < p_aug: randint(0, 7, (B, 2)).
"""
from __future__ import annotations

import numpy as np


class OracleDataset:
    def __init__(self, data: dict, size=None):
        assert 'observations' in data
        self.data = data
        self.size = max(len(v) for v in data.values()) if size is None else size
        self.frame_stack = None
        self.p_aug = None
        self.terminal_locs = np.nonzero(data['terminals'] > 0)[0]
        self.initial_locs = np.concatenate([[0], self.terminal_locs[:-1] + 1])

    @classmethod
    def create_from_initial_dataset(cls, init: dict, size: int):
        n = max(len(v) for v in init.values())
        buf = {}
        for k, v in init.items():
            b = np.zeros((size, *v.shape[1:]), dtype=v.dtype)
            b[: len(v)] = v
            buf[k] = b
        ds = cls(buf, size=None)
        ds.size = ds.pointer = n
        return ds

    def get_random_idxs(self, n):
        return np.random.randint(self.size, size=n)

    def get_subset(self, idxs):
        return {k: v[idxs] for k, v in self.data.items()}

    def sample(self, batch_size, idxs=None):
        if idxs is None:
            idxs = self.get_random_idxs(batch_size)
        batch = self.get_subset(idxs)
        if self.frame_stack is not None:
            initial_state_idxs = self.initial_locs[np.searchsorted(self.initial_locs, idxs, side='right') - 1]
            obs, next_obs = [], []
            for i in reversed(range(self.frame_stack)):
                cur = np.maximum(idxs - i, initial_state_idxs)
                obs.append(self.data['observations'][cur])
                if i != self.frame_stack - 1:
                    next_obs.append(self.data['observations'][cur])
            next_obs.append(self.data['next_observations'][idxs])
            batch['observations'] = np.concatenate(obs, axis=-1)
            batch['next_observations'] = np.concatenate(next_obs, axis=-1)
        if self.p_aug is not None:
            if np.random.rand() < self.p_aug:
                self.augment(batch, ['observations', 'next_observations'])
        return batch

    def augment(self, batch, keys):
        padding = 3
        bs = len(batch[keys[0]])
        crop_froms = np.random.randint(0, 2 * padding + 1, (bs, 2))
        for key in keys:
            arr = batch[key]
            if arr.ndim == 4:
                batch[key] = batched_random_crop(arr, crop_froms, padding)


def batched_random_crop(imgs, crop_froms, padding):
    out = np.empty_like(imgs)
    H, W = imgs.shape[1:3]
    for b in range(len(imgs)):
        padded = np.pad(imgs[b], ((padding, padding), (padding, padding), (0, 0)), mode='edge')
        cy, cx = int(crop_froms[b, 0]), int(crop_froms[b, 1])
        out[b] = padded[cy:cy + H, cx:cx + W]
    return out


def make_synthetic_dataset(n, obs_dim, action_dim, seed=0, episode_len=1000, pixels=False, hw=64):
    """Synthetic OGBench-shaped transitions: datasets are not downloadable offline."""
    rng = np.random.default_rng(seed)
    if pixels:
        obs = rng.integers(0, 256, (n, hw, hw, 3), dtype=np.uint8)
        nobs = rng.integers(0, 256, (n, hw, hw, 3), dtype=np.uint8)
    else:
        obs = rng.standard_normal((n, obs_dim), dtype=np.float32)
        nobs = rng.standard_normal((n, obs_dim), dtype=np.float32)
    act = np.clip(rng.uniform(-1, 1, (n, action_dim)), -1 + 1e-5, 1 - 1e-5).astype(np.float32)
    term = np.zeros(n, np.float32)
    term[episode_len - 1::episode_len] = 1.0
    term[-1] = 1.0
    rew = -(rng.random(n) >= 0.01).astype(np.float32)
    masks = (rew != 0).astype(np.float32)
    return dict(observations=obs, actions=act, next_observations=nobs, rewards=rew, masks=masks, terminals=term)
