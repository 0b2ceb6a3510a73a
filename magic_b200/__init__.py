"""magic_b200 -- B200-native backend for MagIC's radial-loop hot path.

Host-side mirror (Python, over the C ABI of include/magic_sht.h) of the reference interfaces this
backend replaces: `module sht` (sht_native.f90 / shtns.f90), `rIter_t%radialLoop` (rIter.f90) and
`type_mpitransp` (mpi_transpose.f90).  There is no CPU path: every compute call needs the CUDA library
and a CUDA device and fails loudly otherwise.
"""
from .lib import MagicError, load_library  # noqa: F401
from .sht import Sht, grid_sizes  # noqa: F401
from .riter import Params, RadialLoop  # noqa: F401
from .transpose import Transposer, get_blocks  # noqa: F401
from .checkpoint import Checkpoint, read_checkpoint, write_checkpoint  # noqa: F401
