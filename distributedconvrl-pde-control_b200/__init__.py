"""B200-native batched PDE-control hot path (drop-in for DistributedConvRL-PDE-Control's
PDEenv / PDEagent path).  All compute runs in libpdeb200.so (hand-written sm_100a CUDA
behind the C ABI of include/pdeb200.h); this package is the thin host mirror of the
reference's Julia interface.  There is no CPU fallback."""
from . import _lib as lib  # noqa: F401
from ._lib import PdeB200Error  # noqa: F401
from .env import PDEenv  # noqa: F401
from . import setups  # noqa: F401
from . import agent, parallel, checkpoint, hook  # noqa: F401,E402
from .hook import PDEhook  # noqa: F401,E402
