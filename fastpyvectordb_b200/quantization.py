"""Import-compatible alias of the reference module name: ``from fastpyvectordb_b200.quantization import
ScalarQuantizer, BinaryQuantizer, ProductQuantizer`` replaces ``from quantization import ...``."""
from .quantizers import (BinaryQuantizer, DistanceMetric, ProductQuantizer, ScalarQuantizer,  # noqa: F401
                         ScalarQuantizerConfig)
