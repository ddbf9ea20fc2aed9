"""Small invocations of the sweep kernels for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

ge.build()
import tneq_b200 as tb  # noqa: E402

dev = "cuda:0"
for kind, n, K, B in (("merged", 6, 3, 50), ("merged", 5, 2, 37), ("mps", 6, 3, 130), ("tree", 6, 2, 20)):
    torch.manual_seed(1)
    be = tb.BackendFactory.create_backend("b200", device=dev, dtype="float32")
    eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    g = tb.QCTNHelper.generate_example_graph(n=n, graph_type="mps" if kind == "merged" else kind, dim_char=str(K))
    if kind == "merged":
        one = tb.QCTN(g, backend=be)
        g = tb.QCTN.merge(one, one).graph
    q = tb.QCTN(g, backend=be)
    for name in q.cores:
        q.cores_weights[name].requires_grad_(True)
    states = [torch.zeros(K, device=dev) for _ in range(q.nqubits)]
    for s in states:
        s[-1] = 1.0
    mx, _ = eng.generate_data(torch.randn(B, q.nqubits, device=dev), K=K, ret_type="TNTensor")
    vals = eng.contract_with_compiled_strategy(q, states, mx)
    mx, _ = eng.generate_data(torch.randn(B, q.nqubits, device=dev), K=K, ret_type="TNTensor")
    loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, states, mx)
    loss2, grads2 = eng.contract_with_compiled_strategy_for_gradient(q, states, mx, fused=False)
    torch.cuda.synchronize()
    print(kind, n, K, B, float(vals.sum()), float(loss), float(loss2), flush=True)
print("SANITIZE-RUN-DONE")
