#!/usr/bin/env python
"""BASELINE configs[4] / SURVEY 8(d) cfg5: a population of candidate networks evaluated in parallel.

Every candidate is a distinct connected graph (single- and two-layer MPS, tree; 6-16 qubits; edge
rank 2-3; own random cores) that is trained for T steps of the fused step + SGDG on a batch of 512
samples and scored by its final loss -- what tneq_qc/genetic does with one MPI agent per candidate.
Replicas only: candidates are dealt round-robin to the ranks, no traffic until the
final gather of the scores.

    python tools/population_bench.py --candidates 256 --steps 50
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/population_bench.py
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402


def candidates(count):
    fam = []
    for K in (3, 2):
        for n in range(6, 17, 2):
            fam += [("mps", n, K), ("tree", n, K), ("merged", n, K)]
    return [fam[i % len(fam)] + (i,) for i in range(count)]


def evaluate(job):
    """Train and score a list of candidates on one GPU (runs in the rank's process, or in one of its workers)."""
    cands, local, batch, steps, graphs = job
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import tneq_b200 as tb
    out = {}
    for kind, n, K, cid in cands:
        torch.manual_seed(1000 + cid)
        be = tb.BackendFactory.create_backend("b200", device=str(dev), dtype="float32")
        eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
        g = tb.QCTNHelper.generate_example_graph(n=n, graph_type="mps" if kind == "merged" else kind, dim_char=str(K))
        if kind == "merged":
            one = tb.QCTN(g, backend=be)
            g = tb.QCTN.merge(one, one).graph
        q = tb.QCTN(g, backend=be)
        for name in q.cores:
            q.cores_weights[name] = q.cores_weights[name].contiguous().requires_grad_(True)
        states = [torch.zeros(K, device=dev) for _ in range(q.nqubits)]
        for s in states:
            s[-1] = 1.0
        mx, _ = eng.generate_data(torch.randn(batch, q.nqubits, device=dev), K=K, ret_type="TNTensor")
        mx = [tb.TNTensor(m.tensor.contiguous(), m.scale, m.log_scale) for m in mx]
        opt = tb.Optimizer(method="sgdg", learning_rate=0.02, max_iter=steps, engine=eng, momentum=0.9, verbose=False)
        if graphs:       # the fused step replayed from a CUDA graph; the cores ping-pong between two buffers
            eng.enable_cuda_graphs(True)
            opt.opt_state["pingpong"] = True
        loss = None
        for _ in range(steps):
            loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, states, mx)
            opt.step(q, grads)
            opt.iter += 1
        out[cid] = float(loss.detach())
    torch.cuda.synchronize()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=256)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--workers", type=int, default=1,
                    help="worker processes per GPU: a candidate's step is host-bound (~1 ms of Python around ~0.1 ms of "
                         "kernels), so several candidates per GPU are trained concurrently by several processes")
    ap.add_argument("--no-graphs", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    mine = candidates(args.candidates)[rank::world]
    scores = torch.full((args.candidates,), float("nan"), device=dev)
    pool = None
    if args.workers > 1:
        import torch.multiprocessing as mp
        pool = mp.get_context("spawn").Pool(args.workers)
        pool.map(evaluate, [([], local, 1, 0, False)] * args.workers)       # start the workers (CUDA context, imports)
    torch.cuda.synchronize()
    t0 = time.time()
    jobs = [(mine[w::args.workers], local, args.batch, args.steps, not args.no_graphs) for w in range(args.workers)]
    results = pool.map(evaluate, jobs) if pool is not None else [evaluate(jobs[0])]
    for r in results:
        for cid, v in r.items():
            scores[cid] = v
    torch.cuda.synchronize()
    dt = time.time() - t0
    if pool is not None:
        pool.close()
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        scores = torch.nan_to_num(scores, nan=0.0)
        dist.all_reduce(scores)                         # the only exchange: the final gather of the scores
    if rank == 0:
        best = int(torch.argmin(scores).item())
        print(json.dumps({"metric": "candidate networks trained and scored per second", "value": args.candidates / dt,
                          "unit": "candidates/s", "n_gpus": world, "workers_per_gpu": args.workers,
                          "cuda_graphs": not args.no_graphs, "candidates": args.candidates,
                          "steps_per_candidate": args.steps, "batch": args.batch, "seconds": dt,
                          "best_candidate": candidates(args.candidates)[best][:3],
                          "best_loss": float(scores[best]), "scaling": "replicas only"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
