"""`networks.vit` of the reference, served by the sm_100a drop-in (see networks/__init__.py)."""
from hybrid_ctunet_b200.networks.vit import *  # noqa: F401,F403
from hybrid_ctunet_b200.networks import vit as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
