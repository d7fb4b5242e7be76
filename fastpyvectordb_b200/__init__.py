"""B200-native implementation of FastPyVectorDB's data-parallel search hot path.

Drop-in for ``parallel_search.ParallelSearchEngine`` and ``quantization.{Scalar,Binary,Product}Quantizer``;
all distance / selection work runs in hand-written sm_100a CUDA (``libfpv_b200.so``, C-ABI in
``include/fpv_b200.h``).  Importing the package does not need a GPU; using it does, and there is no CPU
fallback.
"""
from .engine import GpuIndex, ParallelSearchEngine, ParallelSearchResult, SearchPipeline
from .quantizers import BinaryQuantizer, DistanceMetric, ProductQuantizer, ScalarQuantizer

__all__ = ["GpuIndex", "ParallelSearchEngine", "ParallelSearchResult", "SearchPipeline", "ScalarQuantizer", "BinaryQuantizer",
           "ProductQuantizer", "DistanceMetric"]
__version__ = "0.1.0"
