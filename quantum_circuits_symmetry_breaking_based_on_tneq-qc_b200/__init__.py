"""tneq_b200 -- B200-native (sm_100a) backend for the QCTN contraction hot path of tneq_qc.

    import tneq_b200
    from tneq_b200 import BackendFactory, EngineSiamese, QCTN, QCTNHelper, Optimizer

    backend = BackendFactory.create_backend("b200", device="cuda:0", dtype="float32")
    engine  = EngineSiamese(backend=backend, strategy_mode="balanced", mx_K=3)
    qctn    = QCTN(QCTNHelper.generate_example_graph(n=16, graph_type="mps", dim_char="3"), backend=backend)

The same names as the reference (tneq_qc.backends / .contractor / .core / .optim)
for the one path this package accelerates; see DESIGN.md and INTEGRATION.md.
Importing the package never touches CUDA; creating the backend does, and raises
if the compiled library or a CUDA device is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"

from .backends.backend_factory import BackendFactory
from .backends.backend_interface import BackendInfo, ComputeBackend
from .backends.backend_b200 import B200Backend
from .contractor.base import ContractionStrategy
from .contractor.compiler import StrategyCompiler
from .contractor.b200_strategy import B200Strategy
from .core.tn_tensor import TNTensor
from .core.qctn import QCTN, QCTNHelper
from .core.engine_siamese import EngineSiamese
from .optim.optimizer import Optimizer

BackendFactory.register_backend("b200", B200Backend)
StrategyCompiler.register_strategy(B200Strategy(), modes=["fast", "balanced", "full"])

__all__ = ["BackendFactory", "BackendInfo", "ComputeBackend", "B200Backend", "ContractionStrategy",
           "StrategyCompiler", "B200Strategy", "TNTensor", "QCTN", "QCTNHelper", "EngineSiamese", "Optimizer"]
