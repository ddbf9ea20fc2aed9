from .tn_tensor import TNTensor
from .qctn import QCTN, QCTNHelper, symbol_of

__all__ = ["TNTensor", "QCTN", "QCTNHelper", "symbol_of"]
