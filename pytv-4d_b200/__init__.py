"""pytv_b200: B200-native drop-in for the GPU path of PyTV-4D (reference pytv/__init__.py:45-63).

    import pytv_b200 as pytv
    tv, G = pytv.tv_GPU.tv_hybrid(img)
    Dx = pytv.tv_operators_GPU.D_hybrid(img)

Only the GPU path exists here: `tv_GPU`, `tv_operators_GPU` (same names and signatures as the reference)
plus `cp` (the README's Chambolle-Pock loop as fused kernels, single- and multi-GPU).  The CUDA library is
loaded on first use and there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib            # noqa: F401  ctypes binding of libpytv_b200.so (lazy load)
from . import tv_operators_GPU  # noqa: F401
from . import tv_GPU          # noqa: F401
from . import cp              # noqa: F401
from . import sharded         # noqa: F401
from .cp import CPSolver, TVProx, cp_denoise, denoise_tv_chambolle, gd_denoise, partition_z  # noqa: F401
from .tv_GPU import TVPlan  # noqa: F401
