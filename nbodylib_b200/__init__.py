"""nbodylib_b200 -- B200-native kd-tree hot path of pelahi/NBodylib (build, kNN, ball search, SPH density,
velocity density, 3D/6D FOF) behind the reference's KDTree interface.  CUDA only; no CPU fallback."""
from .kdtree import KDTree, TPHYS, TPROJ, TVEL, TPHS, TMETRIC, KSPH, KGAUSS, KEPAN, KTH, FOF3D, FOFVEL, FOF6D  # noqa: F401
from ._lib import NbkError, set_option  # noqa: F401


def release_cached_memory(device=-1):
    """Return the library's recycled scratch memory (stream-ordered pool) to the CUDA driver."""
    from . import _lib
    _lib.check(_lib.load().nbk_release_cached_memory(int(device)))
