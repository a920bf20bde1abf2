"""Generates tests/golden/update_goldens.json and sampler_goldens.json.

The reference ships no golden vectors (parity unpinned) and cannot be imported here (no jax/flax/optax), so these are
the repo's own known-answer vectors: inputs are regenerated from fixed seeds (NumPy PCG64 / MT19937 streams are frozen by NumPy's
compatibility policy), outputs come from the fp64 oracle (oracle/fql_oracle.py) resp. the literal sampler transcription
(oracle/sampler_oracle.py).  Run from the repo root:  python tests/golden/make_golden.py
"""
import copy
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fql_oracle as O  # noqa: E402
from oracle import sampler_oracle as SO  # noqa: E402
from tests.helpers import make_case  # noqa: E402

CASES = [  # name, cfg overrides, B, F, A, hidden  -- the four state-based BASELINE shapes + edge cases
    ('cube-single', dict(alpha=300.0), 256, 28, 5, 512),
    ('antmaze-large', dict(q_agg='min', alpha=10.0), 256, 29, 8, 512),
    ('humanoidmaze-medium', dict(discount=0.995, alpha=30.0), 256, 69, 21, 512),
    ('puzzle-4x4', dict(normalize_q_loss=True, alpha=1000.0), 256, 83, 5, 512),
    ('tiny-odd', dict(q_agg='min', alpha=10.0), 37, 11, 5, 64),
    ('one-row', dict(), 1, 4, 2, 64),
]


def leaf_digest(tree):
    out = {}
    for path, v in O.tree_leaves(tree):
        v = np.asarray(v, np.float64)
        flat = v.ravel()
        idx = np.linspace(0, flat.size - 1, min(5, flat.size)).astype(int)
        out['/'.join(path)] = dict(l2=float(np.sqrt((flat ** 2).sum())), sum=float(flat.sum()), absmax=float(np.abs(flat).max()),
                                   samples=[float(flat[i]) for i in idx])
    return out


def main():
    gold = {}
    for name, over, B, F, A, H in CASES:
        cfg, state, batch, noise = make_case(over, B, F, A, seed=sum(map(ord, name)) % 997, hidden=H)
        new_state, info, grads = O.update(copy.deepcopy(state), cfg, batch, noise)
        gold[name] = dict(over=over, B=B, F=F, A=A, hidden=H, seed=sum(map(ord, name)) % 997,
                          info={k: float(v) for k, v in info.items()}, grads=leaf_digest(grads), params=leaf_digest(new_state['params']),
                          mu=leaf_digest(new_state['mu']), nu=leaf_digest(new_state['nu']), count=int(new_state['count']))
        print(name, 'loss', info['critic/critic_loss'], info['actor/actor_loss'])
    json.dump(gold, open(os.path.join(ROOT, 'tests', 'golden', 'update_goldens.json'), 'w'), indent=1)

    sg = {}
    for name, pixels, fs, p_aug, n, ep in [('state-29-8', False, None, None, 5000, 100), ('pixels-fs3-aug', True, 3, 0.5, 300, 37),
                                           ('pixels-plain', True, None, None, 300, 37), ('pixels-fs1-aug1', True, 1, 1.0, 300, 37)]:
        raw = SO.make_synthetic_dataset(n, 29, 8, seed=3, episode_len=ep, pixels=pixels, hw=64)
        ds = SO.OracleDataset.create_from_initial_dataset(raw, size=n + 7)
        ds.frame_stack, ds.p_aug = fs, p_aug
        np.random.seed(11)
        draws = []
        for it in range(4):
            b = ds.sample(32)
            draws.append({k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest() for k, v in b.items()})
        sg[name] = dict(pixels=pixels, frame_stack=fs, p_aug=p_aug, n=n, episode_len=ep, batch=32, draws=draws,
                        rng_after=int(np.random.randint(0, 2 ** 31 - 1)))
    json.dump(sg, open(os.path.join(ROOT, 'tests', 'golden', 'sampler_goldens.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
