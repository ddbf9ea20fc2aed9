#!/usr/bin/env python
"""Single-GPU training of a QCTN through the public API of this package -- the user flow of the
reference's examples/example_train_single_node.py (BASELINE configs[0]) with the `b200` backend.

    python examples/train_single_node.py --num-qubits 16 --dim-char 3 --K 3 --batch-size 512 --num-step 200
    python examples/train_single_node.py --merged --num-qubits 24 --batch-size 16384 --cuda-graphs

Same flags as the reference script (--num-step, --graph-type, --num-qubits, --dim-char, --num-data,
--batch-size, --K, --dtype, --device) plus --merged (two-layer network, QCTN.merge) and --cuda-graphs
(EngineSiamese.enable_cuda_graphs: batches are copied into static device buffers and the training
step is replayed from a CUDA graph).  Note the reference's defect D4: K must equal --dim-char.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tneq_b200 as tb  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--num-step", type=int, default=200)
    ap.add_argument("--save-every", type=int, default=50, help="print the loss every that many steps")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--graph-type", default="mps", choices=["mps", "tree", "wall"])
    ap.add_argument("--merged", action="store_true", help="two layers of the graph (QCTN.merge)")
    ap.add_argument("--num-qubits", type=int, default=16)
    ap.add_argument("--dim-char", default="3")
    ap.add_argument("--num-data", type=int, default=20)
    ap.add_argument("--batch-size", type=int, default=512)
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--learning-rate", type=float, default=1e-2)
    ap.add_argument("--cuda-graphs", action="store_true")
    args = ap.parse_args()

    torch.manual_seed(42)
    backend = tb.BackendFactory.create_backend("b200", device=args.device, dtype=args.dtype)
    engine = tb.EngineSiamese(backend=backend, strategy_mode="balanced", mx_K=args.K)
    dev = backend.backend_info.device
    graph = tb.QCTNHelper.generate_example_graph(n=args.num_qubits, graph_type=args.graph_type, dim_char=args.dim_char)
    if args.merged:
        one = tb.QCTN(graph, backend=backend)
        graph = tb.QCTN.merge(one, one).graph
    qctn = tb.QCTN(graph, backend=backend)
    print(f"backend {backend.get_backend_name()} on {dev}; {qctn.nqubits} qubits, {len(qctn.cores)} cores")
    for name in qctn.cores:                      # as in the reference script: cores are trainable once flagged
        w = qctn.cores_weights[name]
        (w.tensor if isinstance(w, tb.TNTensor) else w).requires_grad_(True)

    tdt = getattr(torch, args.dtype)
    states = [torch.zeros(args.K, device=dev, dtype=tdt) for _ in range(qctn.nqubits)]
    for s in states:
        s[-1] = 1.0
    data = []
    for _ in range(args.num_data):
        x = torch.randn(args.batch_size, qctn.nqubits, device=dev)
        mx, _ = engine.generate_data(x, K=args.K, ret_type="TNTensor")
        data.append(mx)
    if args.cuda_graphs:            # static input buffers: every batch is copied into the same tensors
        engine.enable_cuda_graphs(True)
        static = [torch.empty_like(m.tensor) for m in data[0]]

    opt = tb.Optimizer(method="sgdg", learning_rate=args.learning_rate, max_iter=args.num_step, engine=engine,
                       momentum=0.9, stiefel=True, verbose=False)
    warm = min(10, args.num_step // 2)       # plan compilation / graph capture happen in the first steps
    t0 = time.time()
    for step in range(args.num_step):
        if step == warm:
            torch.cuda.synchronize()
            t0 = time.time()
        mx = data[step % len(data)]
        if args.cuda_graphs:
            for d, m in zip(static, mx):
                d.copy_(m.tensor)
            mx = [tb.TNTensor(d, m.scale, m.log_scale) for d, m in zip(static, mx)]
        loss, grads = engine.contract_with_compiled_strategy_for_gradient(qctn, states, mx)
        opt.step(qctn, grads)
        opt.iter += 1
        if step % args.save_every == 0 or step == args.num_step - 1:
            print(f"step {step:5d}  loss {loss.item():.6f}")
    torch.cuda.synchronize()
    dt = time.time() - t0
    n_timed = args.num_step - warm
    print(f"{n_timed} steps (after {warm} warm-up steps) in {dt:.3f} s: {1e3 * dt / n_timed:.3f} ms/step, "
          f"{n_timed * args.batch_size / dt:,.0f} samples/s (forward + loss + backward + SGDG step, through the public API)")


if __name__ == "__main__":
    main()
