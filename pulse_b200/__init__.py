"""pulse_b200 -- B200-native (sm_100a) implementation of PULSE's per-step rollout / update hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed) calling hand-written CUDA
through the C ABI in include/pulse_b200.h (libpulse_b200.so, built in-tree by pulse_b200.build).
There is no CPU fallback: importing the compute modules without the built library fails loudly.
"""
from . import _lib  # noqa: F401
from ._lib import PulseError  # noqa: F401

__all__ = ["PulseError"]
__version__ = "0.1.0"
