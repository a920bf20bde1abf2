"""fql_b200: B200-native (sm_100a) implementation of the FQL training hot path, `FQLAgent.update` of
zhouzypaul/fql (agents/fql.py), behind the reference's agent API.  See DESIGN.md / INTEGRATION.md."""
from .agent import FQLAgent, INFO_KEYS, NOISE_KEYS  # noqa: F401
from .config import get_config  # noqa: F401

agents = dict(fql=FQLAgent)  # same registry shape as the reference's agents/__init__.py:10-19
