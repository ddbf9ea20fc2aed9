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


class OneShotTimeout(RuntimeError):
    """A peer did not reach the exchange within the timeout (tnq_allreduce_set_timeout_ms)."""


def _precheck(device: torch.device) -> bool:
    """Everything that can fail WITHOUT a collective, so that ranks agree before the first one."""
    try:
        import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
        _lib.load()
        if device.type != "cuda":
            return False
        me = device.index if device.index is not None else torch.cuda.current_device()
        return all(p == me or torch.cuda.can_device_access_peer(me, p) for p in range(torch.cuda.device_count()))
    except Exception:  # noqa: BLE001
        return False


class OneShotAllReduce:
    def __init__(self, nmax: int, device: torch.device, group=None):
        import os
        import torch.distributed._symmetric_memory as symm_mem
        self.lib = _lib.load()
        if os.environ.get("TNQ_ONESHOT_TIMEOUT_S"):
            _lib.check(self.lib.tnq_allreduce_set_timeout_ms(int(float(os.environ["TNQ_ONESHOT_TIMEOUT_S"]) * 1000)))
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
        # ranks first agree on what each can check locally (library, symmetric-memory module, peer
        # access): a rank that cannot take part must not leave its peers inside the rendezvous
        ok = torch.tensor([1.0 if _precheck(device) else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() <= 0:
            return None
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

    def check(self) -> None:
        """Raise OneShotTimeout if any exchange so far gave up waiting for a peer (its output was NaN).
        Synchronises the stream; callers use it when a loss comes back NaN, not on the hot path."""
        torch.cuda.synchronize(self.device)
        words = self.buf.view(torch.int32)[32:35].tolist()
        if words[1] != 0:
            self.buf.view(torch.int32)[33:35].zero_()      # reported once
            raise OneShotTimeout(f"one-shot all-reduce: rank {words[2]} did not arrive in epoch {words[1]} "
                                 f"(rank {self.rank} of {self.world})")
