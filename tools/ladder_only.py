"""One cfg3-shaped training step (24-qubit two-layer merged MPS, K=3) through the C ABI, a few times:
the smallest program that launches the ladder kernels, for ncu (tools/ncu_summary.py reads the report)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import tneq_b200 as tb  # noqa: E402
from oracle import qctn_oracle as oc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
n, K = 24, 3
g1 = tb.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
graph = tb.QCTN.merge(tb.QCTN(g1), tb.QCTN(g1)).graph
names, table, nq = oc.parse_graph(graph)
torch.manual_seed(1234)
cores = oc.random_cores(table)
torch.manual_seed(42)
x = torch.randn(B, nq)
be = tb.BackendFactory.create_backend("b200", device="cuda:0", dtype="float32")
eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
q = tb.QCTN(graph, backend=be)
for k, v in cores.items():
    q.cores_weights[k] = v.cuda().contiguous().requires_grad_(True)
st = [s.cuda() for s in oc.unit_states(nq, K)]
mx, _ = eng.generate_data(x.cuda(), K=K, ret_type="TNTensor")
mx = [tb.TNTensor(m.tensor.contiguous(), m.scale, m.log_scale) for m in mx]
for it in range(4):
    if mode == "train":
        loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, st, mx)
    else:
        with torch.no_grad():
            out = eng.contract_with_compiled_strategy(q, st, mx)
torch.cuda.synchronize()
print("ok", float(loss) if mode == "train" else float(out.sum()))
