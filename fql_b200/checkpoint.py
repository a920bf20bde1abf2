"""Checkpoint wire format of the reference (utils/flax_utils.py:162-202): `params_{epoch}.pkl` is a pickle of
`dict(agent=flax.serialization.to_state_dict(agent))`.  `FQLAgent.state_dict()` produces exactly that nesting
(`rng`, `network/{step, params, opt_state/{0/{count,mu,nu}, 1}}` with the Flax module paths below `params`), as plain
numpy arrays, so a file written here unpickles in the reference without this package and vice versa (jax arrays saved
by the reference are converted with np.asarray on load).  Same function names and arguments as the reference.
"""
from __future__ import annotations

import glob
import os
import pickle

import numpy as np


def _to_numpy(tree):
    if isinstance(tree, dict):
        return {k: _to_numpy(v) for k, v in tree.items()}
    if isinstance(tree, (int, float)) or tree is None:
        return tree
    return np.asarray(tree)


def save_agent(agent, save_dir, epoch):
    """utils/flax_utils.py:162-178."""
    save_dict = dict(agent=_to_numpy(agent.state_dict()))
    save_path = os.path.join(save_dir, f'params_{epoch}.pkl')
    with open(save_path, 'wb') as f:
        pickle.dump(save_dict, f)
    print(f'Saved to {save_path}')


def restore_agent(agent, restore_path, restore_epoch):
    """utils/flax_utils.py:181-202 (`restore_path` is a glob that must match exactly one directory)."""
    candidates = glob.glob(restore_path)
    assert len(candidates) == 1, f'Found {len(candidates)} candidates: {candidates}'
    restore_path = candidates[0] + f'/params_{restore_epoch}.pkl'
    with open(restore_path, 'rb') as f:
        load_dict = pickle.load(f)
    agent.load_state_dict(_to_numpy(load_dict['agent']))
    print(f'Restored from {restore_path}')
    return agent
