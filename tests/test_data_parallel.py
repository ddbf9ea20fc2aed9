"""Host-side logic of the data-parallel path on CPU: two ranks over gloo.

The CUDA contraction itself cannot run here; a stand-in engine returns rank-dependent
losses and gradients so that partitioning, the packed all-reduce (AVG), the core
broadcast and the replicated optimizer step can be checked end to end."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeBackend:
    def optimizer_update(self, params, grads, state, method, hp):
        from tneq_b200.optim import steps
        return steps.optimizer_update(params, grads, state, method, hp)


class _FakeEngine:
    """loss = rank + 1 ; grad of core i = (rank + 1) * (i + 1) everywhere."""

    def __init__(self, rank):
        self.rank, self.backend = rank, _FakeBackend()

    def contract_with_compiled_strategy_for_gradient(self, qctn, circuit_states_list=None, measure_input_list=None):
        scale = float(self.rank + 1) * float(measure_input_list)
        grads = [torch.full_like(qctn.cores_weights[c], scale * (i + 1)) for i, c in enumerate(qctn.cores)]
        return torch.tensor(scale), grads


class _FakeQCTN:
    def __init__(self, rank):
        self.cores = ["a", "b", "c"]
        torch.manual_seed(100 + rank)          # replicas start DIFFERENT on purpose
        self.cores_weights = {c: torch.randn(2, 2, 2, 2) for c in self.cores}


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import tneq_b200  # noqa: F401
    from tneq_b200.distributed import NcclComm, ReduceOp, DataParallelTrainer, TrainingConfig
    comm = NcclComm(backend="gloo")
    assert comm.world_size == world and comm.rank == rank
    # packed list all-reduce, real and complex
    xs = [torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1))),
          torch.full((2,), complex(rank + 1, -(rank + 1)), dtype=torch.complex64)]
    avg = comm.allreduce_list(xs, op=ReduceOp.AVG)
    assert torch.allclose(avg[0], torch.full((3,), 1.5)) and torch.allclose(avg[1], torch.full((2, 2), 15.0))
    assert torch.allclose(avg[2], torch.full((2,), complex(1.5, -1.5), dtype=torch.complex64))
    assert comm.allreduce_scalar(rank + 1.0, ReduceOp.SUM) == 3.0
    # trainer
    qctn, eng = _FakeQCTN(rank), _FakeEngine(rank)
    tr = DataParallelTrainer(eng, qctn, TrainingConfig(optimizer_method="sgd", learning_rate=0.1, log_interval=0,
                                                       tol=0.0), comm=comm)
    data = [{"measure_input_list": float(i + 1)} for i in range(5)]
    part = tr.partition_data(data)
    assert [d["measure_input_list"] for d in part] == ([1.0, 2.0, 3.0] if rank == 0 else [4.0, 5.0])
    tr.sync_model_weights()
    start = {c: qctn.cores_weights[c].clone() for c in qctn.cores}
    loss = tr.train_step(part[0], None)
    # rank0: scale 1*1, rank1: scale 2*4 -> mean loss 4.5 ; mean grad of core i = 4.5 * (i+1)
    assert abs(loss - 4.5) < 1e-6
    for i, c in enumerate(qctn.cores):
        assert torch.allclose(qctn.cores_weights[c], start[c] - 0.1 * 4.5 * (i + 1), atol=1e-6)
    out[rank] = torch.cat([qctn.cores_weights[c].detach().reshape(-1) for c in qctn.cores]).tolist()
    comm.barrier()
    comm.destroy()


def test_two_rank_data_parallel_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert out[0] == out[1], "replicas diverged"
