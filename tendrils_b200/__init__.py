"""tendrils_b200 -- B200-native (sm_100a) implementation of the Tendrils particle step.

The product path is the C ABI in include/tendrils_b200.h (tendrils_b200/lib/libtendrils_b200.so,
built from csrc/ by `python -m tendrils_b200.build`).  This package is the host-side mirror of the
reference's JS classes on top of it.  There is no CPU fallback.
"""
from . import spawn  # noqa: F401
from .flow_line import FlowLine, FlowLines, Line  # noqa: F401
from .optical_flow import OpticalFlow  # noqa: F401
from ._native import TendrilsError, lib_path, load  # noqa: F401
from .tendrils import (INERT, Device, Particles, Shader, Tendrils, Timer, defaults,  # noqa: F401
                       flowShader, initSpawner, logicFrag, shard_columns)

__version__ = "0.1.0"
