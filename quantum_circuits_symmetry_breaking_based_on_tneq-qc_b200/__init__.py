"""tneq_b200: B200-native backend for the QCTN contraction hot path of tneq_qc."""
__version__ = "0.1.0"
