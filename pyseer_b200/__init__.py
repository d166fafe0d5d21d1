"""pyseer_b200 -- B200-native per-variant association engine behind pyseer's interface.

Host code is Python; all per-variant arithmetic runs in hand-written sm_100a CUDA reached
through the C ABI of ``libpyseer_b200.so`` (``include/pyseer_b200.h``) via ctypes.  There is
no CPU fallback: importing the engine without the built library, or creating a context
without a GPU, raises.
"""
__version__ = '0.1.0'

from .classes import Seer, LMM  # noqa: F401
