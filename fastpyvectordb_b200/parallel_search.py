"""Import-compatible alias of the reference module name: ``from fastpyvectordb_b200.parallel_search import
ParallelSearchEngine, ParallelSearchResult`` replaces ``from parallel_search import ...`` (parallel_search.py:59-952;
``ConcurrentHNSWSearcher`` is out of scope)."""
from .engine import (GpuIndex, ParallelSearchEngine, ParallelSearchResult,  # noqa: F401
                     _compute_distances_chunk, _compute_distances_vectorized, _merge_top_k)
from .mmap_store import MemoryMappedVectors, ParallelCollection  # noqa: F401
from .sharded import ShardedSearchEngine, ShardedTopK, shard_bounds  # noqa: F401
