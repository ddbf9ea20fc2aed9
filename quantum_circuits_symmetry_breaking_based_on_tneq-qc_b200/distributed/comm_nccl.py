"""Communicator for data-parallel training: torch.distributed over NCCL (NVLink 5 / NVSwitch).

Replaces, for the ONE collective the hot path needs (gradient + loss averaging),
the reference's CommTorch / CommMPI pair (tneq_qc/distributed/comm/comm_torch.py:102-560,
comm_mpi.py:104-466; Fugaku/MPI path).  Differences that matter:

  * allreduce_list packs all tensors into ONE buffer and issues ONE all-reduce; the
    reference loops `dist.all_reduce` per core tensor (comm_torch.py:510-522), i.e. 46
    latency-bound collectives per step for the 24-qubit two-layer network;
  * AVG is SUM followed by a multiply with 1/world (comm_torch.py:313-316 does the same);
  * the method names the reference's DataParallelTrainer actually calls
    (`allreduce_tensors`, data_parallel.py:204,216 -- missing from every reference comm
    class, SURVEY defect D9) exist here.

One process per GPU; rendezvous from the usual RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT
environment (torchrun).  backend='gloo' is accepted for CPU-side tests of the host logic only.
"""
from __future__ import annotations

import enum
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


class ReduceOp(enum.Enum):
    SUM = "sum"
    AVG = "avg"
    MAX = "max"
    MIN = "min"


_TORCH_OP = {ReduceOp.SUM: dist.ReduceOp.SUM, ReduceOp.AVG: dist.ReduceOp.SUM,
             ReduceOp.MAX: dist.ReduceOp.MAX, ReduceOp.MIN: dist.ReduceOp.MIN}


class NcclComm:
    def __init__(self, backend: str = "nccl", device: Optional[torch.device] = None, init: bool = True):
        self.backend = backend
        self._owns_group = False
        if init and not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            kw = {}
            if backend == "nccl":
                local = int(os.environ.get("LOCAL_RANK", "0"))
                torch.cuda.set_device(local)
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
            self._owns_group = True
        self._initialized = dist.is_initialized()
        self.rank = dist.get_rank() if self._initialized else 0
        self.world_size = dist.get_world_size() if self._initialized else 1
        if device is None:
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if backend == "nccl" \
                else torch.device("cpu")
        self.device = device

    # -- identity ------------------------------------------------------------------
    def get_rank(self) -> int:
        return self.rank

    def get_world_size(self) -> int:
        return self.world_size

    def is_main_process(self) -> bool:
        return self.rank == 0

    def barrier(self):
        if self._initialized:
            dist.barrier()

    # -- collectives ------------------------------------------------------------------
    def allreduce(self, tensor: torch.Tensor, op: ReduceOp = ReduceOp.SUM) -> torch.Tensor:
        out = tensor.clone().contiguous()
        return self.allreduce_inplace(out, op)

    def allreduce_inplace(self, tensor: torch.Tensor, op: ReduceOp = ReduceOp.SUM) -> torch.Tensor:
        if self._initialized and self.world_size > 1:
            if tensor.is_complex():
                dist.all_reduce(torch.view_as_real(tensor), op=_TORCH_OP[op])
            else:
                dist.all_reduce(tensor, op=_TORCH_OP[op])
            if op == ReduceOp.AVG:
                tensor.mul_(1.0 / self.world_size)
        return tensor

    def allreduce_packed(self, flat: torch.Tensor, op: ReduceOp = ReduceOp.AVG) -> torch.Tensor:
        """One collective over an already packed buffer (gradients + loss of one step)."""
        return self.allreduce_inplace(flat, op)

    def allreduce_list(self, tensors: Sequence[torch.Tensor], op: ReduceOp = ReduceOp.AVG) -> List[torch.Tensor]:
        """All tensors in ONE all-reduce: pack (real view), reduce, unpack into new tensors."""
        if not tensors:
            return []
        reals = [torch.view_as_real(t) if t.is_complex() else t for t in tensors]
        flat = torch.cat([r.reshape(-1) for r in reals])
        self.allreduce_inplace(flat, op)
        out, at = [], 0
        for t, r in zip(tensors, reals):
            piece = flat[at:at + r.numel()].reshape(r.shape)
            at += r.numel()
            out.append(torch.view_as_complex(piece.clone()) if t.is_complex() else piece)
        return out

    # names used by the reference's DataParallelTrainer (data_parallel.py:204,216)
    allreduce_tensors = allreduce_list

    def allreduce_scalar(self, value: float, op: ReduceOp = ReduceOp.SUM, device=None) -> float:
        t = torch.tensor([float(value)], dtype=torch.float64, device=device or self.device)
        return float(self.allreduce_inplace(t, op).item())

    def broadcast_tensor(self, tensor: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self._initialized and self.world_size > 1:
            real = torch.view_as_real(tensor) if tensor.is_complex() else tensor
            if real.is_contiguous():
                dist.broadcast(real, src=src)
            else:                                   # e.g. a core made by init_random_core (a transposed view)
                buf = real.detach().contiguous()
                dist.broadcast(buf, src=src)
                with torch.no_grad():
                    real.copy_(buf)
        return tensor

    def broadcast_tensors_packed(self, tensors: Sequence[torch.Tensor], src: int = 0):
        """Broadcast many small tensors (the cores) with one collective; copies back in place."""
        if not (self._initialized and self.world_size > 1) or not tensors:
            return
        reals = [torch.view_as_real(t) if t.is_complex() else t for t in tensors]
        flat = torch.cat([r.detach().reshape(-1) for r in reals])
        dist.broadcast(flat, src=src)
        at = 0
        with torch.no_grad():
            for r in reals:
                r.copy_(flat[at:at + r.numel()].reshape(r.shape))
                at += r.numel()

    def broadcast_object(self, obj, src: int = 0):
        if not (self._initialized and self.world_size > 1):
            return obj
        box = [obj if self.rank == src else None]
        dist.broadcast_object_list(box, src=src)
        return box[0]

    def destroy(self):
        if self._owns_group and dist.is_initialized():
            dist.destroy_process_group()
        self._initialized = False
