"""State containers mirroring sgmcmcjax/types.py:14-34."""
from __future__ import annotations

from typing import Any, NamedTuple, Optional


class DiffusionState(NamedTuple):
    """types.py:14-22.  packed_state holds the flat per-chain device buffers
    {"theta", "aux1", "aux2", "alpha"} instead of per-leaf jax arrays."""
    packed_state: Any
    tree_def: Any
    subtree_defs: Any = None


class SVRGState(NamedTuple):
    """types.py:25-27."""
    centering_value: Any = None
    fb_grad_center: Any = None


class SamplerState(NamedTuple):
    """types.py:30-34."""
    diffusion_state: DiffusionState
    param_grad: Any
    svrg_state: SVRGState = SVRGState()
    grad_info: Optional[Any] = None
