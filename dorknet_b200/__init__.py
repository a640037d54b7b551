"""dorknet_b200: a from-scratch, B200-native (sm_100a) implementation of Dorknet's CNN training hot
path behind Dorknet's own layer API.  Host code is Python; every computation is a hand-written CUDA
kernel reached through the C ABI in include/dorknet_b200.h.  There is no CPU fallback."""
from . import runtime  # noqa: F401
from ._lib import api, launch_count, load  # noqa: F401
from .array import DeviceArray, DeviceScalar, asarray, asnumpy, empty, zeros  # noqa: F401

__version__ = "0.1.0"
