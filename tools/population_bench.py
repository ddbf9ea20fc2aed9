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


class Candidate:
    """One candidate network: its own engine, cores, batch, optimizer -- and its own CUDA stream, so that the small,
    latency-bound kernels of several candidates overlap on the GPU."""

    _streams = {}       # (device, slot) -> stream: reused by successive candidates, so that the caching allocator's
                        # per-stream pools are reused too (a fresh stream per candidate means cudaMalloc per candidate)

    def __init__(self, spec, dev, batch, steps, graphs, slot=0):
        import tneq_b200 as tb
        kind, n, K, cid = spec
        if (dev, slot) not in Candidate._streams:
            Candidate._streams[(dev, slot)] = torch.cuda.Stream(dev) if slot > 0 else torch.cuda.current_stream(dev)
        self.cid, self.stream, self.loss = cid, Candidate._streams[(dev, slot)], None
        with torch.cuda.stream(self.stream):
            torch.manual_seed(1000 + cid)
            be = tb.BackendFactory.create_backend("b200", device=str(dev), dtype="float32")
            self.eng = eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
            g = tb.QCTNHelper.generate_example_graph(n=n, graph_type="mps" if kind == "merged" else kind, dim_char=str(K))
            if kind == "merged":
                one = tb.QCTN(g, backend=be)
                g = tb.QCTN.merge(one, one).graph
            self.q = q = tb.QCTN(g, backend=be)
            for name in q.cores:
                q.cores_weights[name] = q.cores_weights[name].contiguous().requires_grad_(True)
            self.states = [torch.zeros(K, device=dev) for _ in range(q.nqubits)]
            for s in self.states:
                s[-1] = 1.0
            mx, _ = eng.generate_data(torch.randn(batch, q.nqubits, device=dev), K=K, ret_type="TNTensor")
            self.mx = [tb.TNTensor(m.tensor.contiguous(), m.scale, m.log_scale) for m in mx]
            self.opt = tb.Optimizer(method="sgdg", learning_rate=0.02, max_iter=steps, engine=eng, momentum=0.9, verbose=False)
            if graphs:       # the fused step replayed from a CUDA graph; the cores ping-pong between two buffers
                eng.enable_cuda_graphs(True)
                self.opt.opt_state["pingpong"] = True

    def step(self):
        with torch.cuda.stream(self.stream):
            self.loss, grads = self.eng.contract_with_compiled_strategy_for_gradient(self.q, self.states, self.mx)
            self.opt.step(self.q, grads)
            self.opt.iter += 1


def evaluate(job):
    """Train and score a list of candidates on one GPU (runs in the rank's process, or in one of its workers):
    `streams` candidates at a time, their steps issued round-robin on their own CUDA streams."""
    cands, local, batch, steps, graphs, streams = job
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    out = {}
    for i in range(0, len(cands), max(1, streams)):
        group = [Candidate(spec, dev, batch, steps, graphs, slot) for slot, spec in enumerate(cands[i:i + max(1, streams)])]
        for _ in range(steps):
            for c in group:
                c.step()
        torch.cuda.synchronize(dev)
        for c in group:
            out[c.cid] = float(c.loss.detach())
    torch.cuda.synchronize()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=256)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--workers", type=int, default=0,
                    help="worker processes per GPU: a candidate's step is host-bound (~1 ms of Python around ~0.1 ms of "
                         "kernels), so several candidates per GPU are trained concurrently by several processes")
    ap.add_argument("--streams", type=int, default=1,
                    help="candidates trained concurrently by ONE process, each on its own CUDA stream (their kernels are small "
                         "and latency-bound: they can overlap on the GPU; measured on B200 this is NOT faster than worker processes -- "
                         "19.5 candidates/s with 8 streams vs 19.7 with one -- the step is bound by Python, not by the device)")
    ap.add_argument("--no-graphs", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if args.workers <= 0:       # one process per host core, shared by the ranks of the node (a candidate's step is host-bound)
        args.workers = max(1, min(8, (os.cpu_count() or 1) // max(1, world)))
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
    # untimed warm-up in every process that will evaluate candidates: CUDA context, library and kernel module loading,
    # the cuSOLVER handle behind the QR initialisation of the cores -- seconds of one-time cost that a search over
    # thousands of candidates pays once
    warm = ([c[:3] + (10_000 + i,) for i, c in enumerate(candidates(3))], local, args.batch, 4, not args.no_graphs, args.streams)
    if args.workers > 1:
        import torch.multiprocessing as mp
        pool = mp.get_context("spawn").Pool(args.workers)
        pool.map(evaluate, [warm] * args.workers)
    else:
        evaluate(warm)
    torch.cuda.synchronize()
    t0 = time.time()
    jobs = [(mine[w::args.workers], local, args.batch, args.steps, not args.no_graphs, args.streams) for w in range(args.workers)]
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
                          "unit": "candidates/s", "n_gpus": world, "workers_per_gpu": args.workers, "streams_per_worker": args.streams,
                          "cuda_graphs": not args.no_graphs, "candidates": args.candidates,
                          "steps_per_candidate": args.steps, "batch": args.batch, "seconds": dt,
                          "best_candidate": candidates(args.candidates)[best][:3],
                          "best_loss": float(scores[best]), "scaling": "replicas only"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
