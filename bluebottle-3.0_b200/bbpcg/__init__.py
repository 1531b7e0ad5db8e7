"""bbpcg -- host-side mirror of Bluebottle's pressure-Poisson entry points on top of the
B200-native C-ABI library (../lib/libbbpcg.so, built from ../csrc by __graft_entry__.build()).

PyTorch is used only as plumbing: device memory, streams and torch.distributed for the
multi-process launch.  There is no CPU fallback: importing works anywhere, but every compute
call needs the CUDA library and a GPU and fails loudly otherwise.
"""
from . import grid, synth  # noqa: F401
from .lib import load_library, LibraryMissing  # noqa: F401
from .solver import PoissonSolver, SolveResult, Decomposition, read_restart, restart_path  # noqa: F401
