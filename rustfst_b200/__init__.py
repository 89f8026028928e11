"""rustfst_b200 — B200-native compose / shortest-path behind the rustfst-ffi C-ABI.

Host-side mirror of the rustfst Python package (rustfst-python/rustfst) for the compose / shortest-path slice:
same class and function names, backed by librustfst_b200.so (hand-written sm_100a CUDA kernels).
"""
from .ffi import check_ffi_error, device_count, lib  # noqa: F401
from .fst import TR_DTYPE, ConstFst, Tr, TrsIterator, VectorFst, weight_one, weight_zero  # noqa: F401
from .algorithms import (AcceptorBatch, ComposeConfig, ComposeFilter, DeviceFst, MatcherConfig, MatcherRewriteMode,  # noqa: F401
                         PackedBatch, ShortestPathConfig, compose, compose_batch, compose_batch_packed, compose_with_config, compose_with_stats,
                         device_compose, device_shortest_path, shortestpath, shortestpath_with_config,
                         shortestpath_with_stats)
from . import props  # noqa: F401
