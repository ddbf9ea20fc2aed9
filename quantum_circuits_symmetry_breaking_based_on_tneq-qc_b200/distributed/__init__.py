from .comm_nccl import NcclComm, ReduceOp
from .data_parallel import DataParallelTrainer, TrainingConfig, TrainingStats

__all__ = ["NcclComm", "ReduceOp", "DataParallelTrainer", "TrainingConfig", "TrainingStats"]
