"""Data-parallel training on two GPUs through the real CUDA path (ladder kernel + one-shot NVLink
all-reduce + batched SGDG step) against a single-process emulation of the same schedule.

DataParallelTrainer semantics (reference data_parallel.py:266-307): every rank computes the
mean-loss gradient of ITS batch, the gradients are averaged over ranks, every rank applies the same
optimizer step.  So after every step the cores must equal those of one process that averages the
gradients of the two batches itself -- and the two replicas must be bit-identical to each other.

Needs two visible GPUs (skipped on the one-GPU tier; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp2.py -m gpu`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, random, sys
sys.path.insert(0, os.environ["TNQ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["TNQ_ROOT"], "tests"))
import torch, torch.distributed as dist
import tneq_b200 as tb
from tneq_b200.distributed import NcclComm, DataParallelTrainer, TrainingConfig
from oracle import qctn_oracle as oc

rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
comm = NcclComm(backend="nccl", device=dev)
n, K, B, STEPS = 6, 3, 96, 4
g1 = tb.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
graph = tb.QCTN.merge(tb.QCTN(g1), tb.QCTN(g1)).graph
names, table, nq = oc.parse_graph(graph)
torch.manual_seed(5)                      # identical on both ranks: cores, data
cores0 = oc.random_cores(table)
xs = [0.7 * torch.randn(B, nq) for _ in range(2)]


def fresh(device):
    be = tb.BackendFactory.create_backend("b200", device=str(device), dtype="float32")
    eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    q = tb.QCTN(graph, backend=be)
    for k, v in cores0.items():
        q.cores_weights[k] = v.to(device).clone().requires_grad_(True)
    st = [s.to(device) for s in oc.unit_states(nq, K)]
    data = []
    for x in xs:
        mx, _ = eng.generate_data(x.to(device), K=K, ret_type="tensor")
        data.append({"measure_input_list": [m.contiguous() for m in mx]})
    return eng, q, st, data


cfg = TrainingConfig(optimizer_method="sgdg", learning_rate=0.05, momentum=0.9, log_interval=0, tol=0.0, seed=9)
eng, q, st, data = fresh(dev)
tr = DataParallelTrainer(eng, q, cfg, comm=comm)
mine = tr.partition_data(data)
assert len(mine) == 1
tr.sync_model_weights()
losses = [tr.train_step(mine[0], st) for _ in range(STEPS)]
assert tr._oneshot is not None, "the one-shot NVLink all-reduce was expected to be active"
got = torch.cat([q.cores_weights[c].detach().reshape(-1) for c in q.cores])

# replicas bit-identical
both = [torch.empty_like(got) for _ in range(2)]
dist.all_gather(both, got)
assert torch.equal(both[0], both[1]), "replicas diverged"

# single-process emulation of the same schedule on this rank's GPU
eng2, q2, st2, data2 = fresh(dev)
opt = tb.Optimizer(method="sgdg", learning_rate=0.05, engine=eng2, momentum=0.9, stiefel=True, verbose=False)
random.seed(cfg.seed)
ref_losses = []
for _ in range(STEPS):
    ls, gs = [], []
    for d in data2:
        l, g = eng2.contract_with_compiled_strategy_for_gradient(q2, st2, **d)
        ls.append(float(l)); gs.append([x.clone() for x in g])
    avg = [(a + b) / 2 for a, b in zip(*gs)]
    opt.step(q2, avg)
    opt.iter += 1
    ref_losses.append(sum(ls) / 2)
want = torch.cat([q2.cores_weights[c].detach().reshape(-1) for c in q2.cores])
err = ((got - want).abs().max() / want.abs().max()).item()
assert err < 2e-5, f"data-parallel cores differ from the single-process schedule: {err:.2e}"
for a, b in zip(losses, ref_losses):
    assert abs(a - b) <= 1e-5 * abs(b), (a, b)
assert losses[-1] < losses[0]

# the same schedule with ONE CUDA-graph launch per step (contraction + exchange captured together, cores
# ping-ponging between two buffers): same numbers, and the graph really replays
eng3, q3, st3, data3 = fresh(dev)
tr3 = DataParallelTrainer(eng3, q3, cfg, comm=comm)
tr3.enable_cuda_graphs(True)
mine3 = tr3.partition_data(data3)
tr3.sync_model_weights()
losses3 = [tr3.train_step(mine3[0], st3) for _ in range(STEPS + 6)]
fn3 = eng3._compiled(q3, st3, mine3[0]["measure_input_list"], True, "symmetric")
assert fn3.graph_stats["replays"] >= 2, fn3.graph_stats
for a, b in zip(losses3, losses):
    assert abs(a - b) <= 2e-5 * abs(b), (a, b)
tr3.enable_cuda_graphs(False)
comm.barrier()
comm.destroy()
print("DP2-OK", rank, err)
'''


def test_two_gpu_training_matches_single_process_schedule(built_lib, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "dp2_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TNQ_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("DP2-OK") == 2


WORKER_OVERLAP = r'''
import os, sys
os.environ["TNQ_FORCE_GEMM_PATH"] = "1"          # the route bond 64-128 takes, at a size that runs in seconds
sys.path.insert(0, os.environ["TNQ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["TNQ_ROOT"], "tests"))
import torch, torch.distributed as dist
import tneq_b200 as tb
from tneq_b200.distributed import NcclComm, DataParallelTrainer, TrainingConfig
from oracle import qctn_oracle as oc

rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
comm = NcclComm(backend="nccl", device=dev)
n, K, B, STEPS = 5, 8, 12, 3
graph = tb.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
names, table, nq = oc.parse_graph(graph)
torch.manual_seed(11)
cores0 = oc.random_cores(table, torch.complex64)
mats = []
for r in range(2):                                # one batch per rank, identical on both ranks
    ms = []
    for q_ in range(nq):
        v = torch.randn(B, K, dtype=torch.complex64)
        ms.append(torch.einsum("bi,bj->bij", v.conj(), v) / K)
    mats.append(ms)


def run(overlap):
    be = tb.BackendFactory.create_backend("b200", device=str(dev), dtype="complex64")
    eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    q = tb.QCTN(graph, backend=be)
    for k, v in cores0.items():
        q.cores_weights[k] = v.to(dev).clone().requires_grad_(True)
    st = [s.to(dev).to(torch.complex64) for s in oc.unit_states(nq, K)]
    data = [{"measure_input_list": [m.to(dev).contiguous() for m in ms]} for ms in mats]
    cfg = TrainingConfig(optimizer_method="sgd", learning_rate=0.05, log_interval=0, tol=0.0, seed=9, overlap_exchange=overlap)
    tr = DataParallelTrainer(eng, q, cfg, comm=comm)
    mine = tr.partition_data(data)
    tr.sync_model_weights()
    losses = [tr.train_step(mine[0], st) for _ in range(STEPS)]
    fired = tr._ovl["fired"] if tr._ovl is not None else 0
    return losses, torch.cat([torch.view_as_real(q.cores_weights[c].detach()).reshape(-1) for c in q.cores]), fired, len(q.cores)


la, ca, fired, ncores = run(True)
assert fired == ncores, (fired, ncores)           # every core was reduced from inside the reverse sweep
lb, cb, fired_b, _ = run(False)
assert fired_b == 0
err = ((ca - cb).abs().max() / cb.abs().max()).item()
assert err < 1e-6, f"overlapped per-core exchange differs from the packed exchange: {err:.2e}"
for a, b in zip(la, lb):
    assert abs(a - b) <= 1e-6 * abs(b), (a, b)
both = [torch.empty_like(ca) for _ in range(2)]
dist.all_gather(both, ca)
assert torch.equal(both[0], both[1]), "replicas diverged"
comm.barrier()
comm.destroy()
print("DP2-OVERLAP-OK", rank, err)
'''


def test_two_gpu_large_bond_exchange_overlaps_the_reverse_sweep(built_lib, tmp_path):
    """DataParallelTrainer on the large-bond (GEMM) route: every core's all-reduce is started from inside the reverse
    sweep (compute_fn.set_grad_ready_hook) -- same cores and losses as the packed exchange after the step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "dp2_overlap_worker.py"
    script.write_text(WORKER_OVERLAP)
    env = dict(os.environ, TNQ_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29539", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
