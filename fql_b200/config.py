"""get_config() of agents/fql.py:249-270, without ml_collections (not installed here; used if importable)."""
from __future__ import annotations


class ConfigDict(dict):
    """Minimal stand-in for ml_collections.ConfigDict: item and attribute access, unlocked."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self):
        return dict(self)


def get_config():
    cfg = dict(
        agent_name='fql',
        ob_dims=None,              # set by create()
        action_dim=None,           # set by create()
        lr=3e-4,
        batch_size=256,
        actor_hidden_dims=(512, 512, 512, 512),
        value_hidden_dims=(512, 512, 512, 512),
        layer_norm=True,
        actor_layer_norm=False,
        discount=0.99,
        tau=0.005,
        q_agg='mean',
        alpha=300.0,
        flow_steps=10,
        normalize_q_loss=False,
        encoder=None,
        # keys this fork's main.py reads every iteration (main.py:198); None = not used by FQL
        actor_start_steps=None,
        critic_train_steps=None,
    )
    try:
        import ml_collections
        c = ml_collections.ConfigDict(cfg)
        return c
    except ImportError:
        return ConfigDict(cfg)
