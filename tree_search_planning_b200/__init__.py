"""B200-native batched MuZero tree search (planning hot path of Disiok/tree-search-planning).

Public surface:
  MCTS            drop-in for self_play.MCTS (+ stochastic / risk-sensitive siblings), plus run_batch
  Engine          ctypes binding of the C ABI (include/tsp.h) over libtsp_b200.so
  get_config      the reference game-file hyper-parameters as attribute bags
  selfplay        BatchedSelfPlay (the per-step loop of SelfPlay.play_game for B games at once, root I/O formats on
                  the device) and reanalyse_values (imported lazily: needs torch for device buffers)
"""
from .configs import SearchConfig, get_config, variant_of, observation_dim, PRESET_NAMES  # noqa: F401
from .engine import Engine, TspError, load_library, ABI_SYMBOLS  # noqa: F401
from .mcts import MCTS, BatchResult  # noqa: F401

__version__ = "0.1.0"
