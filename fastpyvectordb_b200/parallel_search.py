"""Import-compatible alias of the reference module name: ``from fastpyvectordb_b200.parallel_search import
ParallelSearchEngine, ParallelSearchResult`` replaces ``from parallel_search import ...``."""
from .engine import GpuIndex, ParallelSearchEngine, ParallelSearchResult  # noqa: F401
