"""One-shot NVLink all-reduce of the packed gradient + loss buffer (csrc/tnq_allreduce.cu).

The buffers every rank publishes into are symmetric memory (torch.distributed._symmetric_memory:
allocated once, mapped into every peer over NVLink / NVSwitch); the reduction itself is this
repository's kernel `tnq_allreduce_oneshot` -- PyTorch only provides the allocation and the
pointer exchange.  If symmetric memory cannot be set up (single GPU, no peer access) `create`
returns None and the caller keeps the NCCL all-reduce.

Replaces the per-core blocking collectives of the reference's
DataParallelTrainer.sync_gradients (tneq_qc/distributed/parallel/data_parallel.py:194-204).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .. import _lib


class OneShotAllReduce:
    def __init__(self, nmax: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.lib = _lib.load()
        self.nmax, self.device = int(nmax), device
        group = group if group is not None else dist.group.WORLD
        words = int(self.lib.tnq_allreduce_oneshot_words(self.nmax))
        self.buf = symm_mem.empty(words, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        self.rank, self.world = int(self.handle.rank), int(self.handle.world_size)
        self.peers_dev = int(self.handle.buffer_ptrs_dev)

    @classmethod
    def create(cls, nmax: int, device: torch.device, group=None) -> Optional["OneShotAllReduce"]:
        if not (dist.is_available() and dist.is_initialized()) or not 2 <= dist.get_world_size() <= 16:
            return None
        if dist.get_backend() != "nccl":
            return None
        ok = torch.ones(1, device=device)
        try:
            obj = cls(nmax, device, group)
        except Exception as exc:  # noqa: BLE001 -- any set-up failure means "use NCCL"
            obj = None
            ok.zero_()
            print(f"[tneq_b200] rank {dist.get_rank()}: symmetric memory unavailable ({type(exc).__name__}: {exc}); "
                  "using the NCCL all-reduce", flush=True)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # all ranks or none
        return obj if ok.item() > 0 else None

    def mean(self, flat: torch.Tensor, extra: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(flat ++ extra) averaged over the ranks, as a new tensor of flat.numel() + extra.numel()."""
        na = flat.numel()
        nb = extra.numel() if extra is not None else 0
        if flat.dtype != torch.float32 or not flat.is_contiguous() or (extra is not None and extra.dtype != torch.float32):
            raise ValueError("one-shot all-reduce takes contiguous float32 buffers")
        out = torch.empty(na + nb, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.tnq_allreduce_oneshot(self.peers_dev, self.rank, self.world, self.nmax, flat.data_ptr(), na,
                                                      extra.data_ptr() if extra is not None else None, nb, out.data_ptr(),
                                                      1.0 / self.world, torch.cuda.current_stream(self.device).cuda_stream))
        return out
