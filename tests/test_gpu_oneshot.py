"""The one-shot NVLink all-reduce kernel (csrc/tnq_allreduce.cu) against NCCL, on every visible GPU
(two or more), and its behaviour when a peer is late.

Needs at least two visible GPUs (the round-end GPU tier has one: skipped there; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_oneshot.py -m gpu`, logs kept in profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["TNQ_ROOT"])
import torch, torch.distributed as dist
import tneq_b200
from tneq_b200.distributed.oneshot import OneShotAllReduce
rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n = 3726
red = OneShotAllReduce.create(n + 16, dev)
assert red is not None, "symmetric memory could not be set up"
for it in range(50):                       # many epochs: exercises the double-buffered slots
    torch.manual_seed(100 * it + rank)
    flat = torch.randn(n, device=dev)
    loss = torch.randn(1, device=dev)
    got = red.mean(flat, loss)
    want = torch.cat([flat, loss])
    dist.all_reduce(want)
    want /= dist.get_world_size()
    # both sum in float32; the one-shot kernel sums in rank order on every rank
    assert (got - want).abs().max().item() <= 1e-6 * want.abs().max().item(), it
    gathered = [torch.empty_like(got) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, got)
    assert all(torch.equal(g, gathered[0]) for g in gathered)      # bit-identical on all ranks
# a late peer: rank 1 arrives after the timeout.  The early ranks get NaN + OneShotTimeout (no trap,
# the context survives), the late rank completes with everyone's data, and the next exchange is
# back in step on all ranks.
import time
from tneq_b200.distributed.oneshot import OneShotTimeout
from tneq_b200 import _lib
_lib.check(_lib.load().tnq_allreduce_set_timeout_ms(300))
dist.barrier(); torch.cuda.synchronize()
flat = torch.full((n,), float(rank + 1), device=dev)
if rank == 1:
    time.sleep(2.0)
got = red.mean(flat, None)
torch.cuda.synchronize()
if rank == 1:
    assert torch.isfinite(got).all() and abs(got[0].item() - (dist.get_world_size() + 1) / 2) < 1e-6
    red.check()
else:
    assert torch.isnan(got).all(), "an exchange that gave up must not publish numbers"
    try:
        red.check()
        raise AssertionError("check() did not raise")
    except OneShotTimeout as exc:
        assert "rank 1" in str(exc)
_lib.check(_lib.load().tnq_allreduce_set_timeout_ms(600000))
dist.barrier()
got = red.mean(flat, None)
torch.cuda.synchronize()
assert abs(got[5].item() - (dist.get_world_size() + 1) / 2) < 1e-6
dist.barrier()
dist.destroy_process_group()
print("ONESHOT-OK", rank)
'''


def test_oneshot_allreduce_matches_nccl(built_lib, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TNQ_ROOT=ROOT)
    world = min(torch.cuda.device_count(), 8)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ONESHOT-OK") == world
