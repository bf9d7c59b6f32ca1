"""bout-dev_b200: B200-native implementation of the BOUT++ physics-model RHS hot path.

The product is the CUDA library ``csrc/libbout_b200.so`` (C-ABI in ``include/bout_b200.h``) and
the C++ header layer in ``host/`` that presents BOUT++'s own API on top of it.  This Python
package is only the harness binding used by tests/ and bench.py (``capi``): it loads the shared
library through ctypes and passes torch CUDA tensors' device pointers.  There is no CPU fallback:
importing ``capi`` without the built library raises.
"""
from . import capi  # noqa: F401
from . import mesh  # noqa: F401
from . import models  # noqa: F401
from . import host  # noqa: F401
from . import elm_setup  # noqa: F401
