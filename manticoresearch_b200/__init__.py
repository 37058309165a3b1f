"""manticoresearch_b200 -- B200-native full-text query hot path behind Manticore's operator surface.

The product is libmgpu.so (hand-written sm_100a CUDA + C++ host, C ABI in include/mgpu.h).  This
package is a thin ctypes mirror of that ABI for the tests and the benchmark; it contains no
search logic and no CPU fallback: without the built library, or without a GPU, calls fail loudly.
"""
from .mgpu import (  # noqa: F401
    Index, Query, Node, Keyword, kw, AND, OR, ANDNOT, MAYBE, PHRASE, PROXIMITY,
    SortKey, Filter, MgpuError, lib, load_library, build_index, build_synthetic, SynthParams,
    RANK_PROXIMITY_BM25, RANK_BM25, RANK_NONE, RANK_WORDCOUNT,
    KEYPART_ROWID, KEYPART_WEIGHT, KEYPART_INT, FILTER_RANGE, FILTER_VALUES,
    MGPU_OK, MGPU_E_UNSUPPORTED, MGPU_E_NO_DEVICE,
)
