"""B200-native AlphaZero self-play engine behind the alpha-zero-general plugin surface.

Host-side mirror of the reference's interfaces for the self-play hot path:
  game.SplendorGame   <->  splendor/SplendorGame.py (Game.py)
  nnet.NNetWrapper    <->  splendor/NNet.py / GenericNNetWrapper.py (predict half); nnet.SantoriniNNetWrapper <-> santorini/NNet.py
  mcts.MCTS           <->  MCTS.py
  coach.Coach         <->  Coach.py (executeEpisode / executeEpisodes)
All compute goes through the C ABI in include/azg.h (csrc/libazg_b200.so, hand-written sm_100a CUDA).
There is no CPU fallback: importing works anywhere, computing needs the built library and a GPU.
"""
from . import lib  # noqa: F401
from .game import SplendorGame, SantoriniGame, AbaloneGame, AzulGame, CudaGame  # noqa: F401
from .nnet import NNetWrapper, SantoriniNNetWrapper, AbaloneNNetWrapper, AzulNNetWrapper, V84_TENSOR_ORDER, V80_TENSOR_ORDER, V89_TENSOR_ORDER, V21_TENSOR_ORDER  # noqa: F401
from .mcts import MCTS  # noqa: F401
from .coach import Coach  # noqa: F401
from .utils import dotdict  # noqa: F401

__all__ = ['lib', 'SplendorGame', 'SantoriniGame', 'AbaloneGame', 'AzulGame', 'CudaGame', 'NNetWrapper', 'SantoriniNNetWrapper', 'AbaloneNNetWrapper', 'AzulNNetWrapper', 'V84_TENSOR_ORDER', 'MCTS', 'Coach', 'dotdict', 'V80_TENSOR_ORDER', 'V89_TENSOR_ORDER', 'V21_TENSOR_ORDER']
