"""B200-native D2Q9-BGK lattice-Boltzmann time step behind the d2q9-bgk surface.

The product is native: ``csrc/`` (sm_100a CUDA kernels + the C-ABI of ``include/lbm.h``,
built to ``liblbm_b200.so``) and ``host/`` (the C program ``d2q9-bgk``).  The Python here is
the same surface for ``tests/`` and ``bench.py``: ``decks`` (file formats, host maths) and
``cabi`` (ctypes binding).  Import as ``opencl_lattice_boltzmann_b200`` (a symlink to this
directory, whose name has a hyphen).
"""
from . import decks  # noqa: F401
from . import cabi  # noqa: F401
from . import ring  # noqa: F401
from .build import build_all  # noqa: F401

__all__ = ["decks", "cabi", "ring", "build_all"]
