"""spsph -- Python mirror of the reference's driver-side interface (Init_sph / time_integration / OutputRes)
over the C-ABI of include/spsph.h. The compute path lives in libspsph_cuda.so (hand-written sm_100a CUDA);
there is no CPU fallback: constructing an Engine without the CUDA library or without a GPU raises."""
from . import _abi  # noqa: F401
from . import checkpoint  # noqa: F401
from . import dist  # noqa: F401
from .engine import Engine, cuda_lib, dist_unique_id  # noqa: F401
from .problem import Problem, load  # noqa: F401
