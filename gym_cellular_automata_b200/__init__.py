"""gym_cellular_automata_b200 -- B200-native batched forest-fire CA environment step.

Scope: the environment-step hot path of frasermince/gym-cellular-automata's
advanced bulldozer env behind the reference's Operator / CAEnv API.  Hand-written sm_100a CUDA in
csrc/, reached through the C ABI of include/gca.h (libgca.so) via ctypes; torch tensors are the
zero-copy device buffers.  No CPU fallback: importing the env classes works anywhere, running them
needs the built library and a GPU."""
from .ca_env import CAEnv
from .grid_space import GridSpace
from .operator import Operator
from . import spaces

__version__ = "0.1.0"
__all__ = ["CAEnv", "GridSpace", "Operator", "spaces"]
