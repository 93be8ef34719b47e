"""velesdb_b200 -- B200-native (sm_100a) implementation of VelesDB's vector-search hot path.

The product is ``csrc/libveles_b200.so`` (hand-written CUDA behind the C ABI of
``include/veles_b200.h``).  The Python modules mirror the reference's host-side wrapper types
(``HnswIndex``, ``Bm25Index``, fusion) so the path can be driven and tested without Rust.
There is no CPU fallback anywhere in this package.
"""
from . import _native
from ._native import VelesError
from .bm25 import Bm25Index, Bm25Params, Bm25Snapshot, tokenize
from .dual_precision import DualPrecisionConfig, DualPrecisionHnsw
from .fusion import (FusionError, FusionStrategy, hybrid_search, hybrid_search_batch, multi_query_search, overfetch_k, rrf_hybrid_batch,
                     search_with_filter)
from .index import (DeviceSnapshot, DimensionMismatch, DistanceMetric, HnswIndex, HnswParams, SearchQuality,
                    VacuumError, distance_pairs, multi_entry_probes)

__all__ = ["DualPrecisionConfig", "DualPrecisionHnsw", "Bm25Index", "Bm25Params", "Bm25Snapshot", "FusionError", "FusionStrategy", "hybrid_search", "hybrid_search_batch", "multi_query_search", "overfetch_k", "search_with_filter",
           "rrf_hybrid_batch", "tokenize", "DeviceSnapshot", "DimensionMismatch", "DistanceMetric", "HnswIndex", "HnswParams", "SearchQuality",
           "VacuumError", "VelesError", "distance_pairs", "multi_entry_probes", "_native"]
