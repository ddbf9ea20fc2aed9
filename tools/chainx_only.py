"""cfg2-large-shaped forward from x (16-qubit MPS, K=3, batch 2^20) through EngineSiamese.contract_from_x, a few
times: the smallest program that launches tnq_hermite_scale_kernel + tnq_chain_fwd2x_kernel, for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import tneq_b200 as tb  # noqa: E402
from oracle import qctn_oracle as oc  # noqa: E402

B, n, K = 1 << 20, 16, 3
graph = tb.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
names, table, nq = oc.parse_graph(graph)
torch.manual_seed(1234)
cores = oc.random_cores(table)
be = tb.BackendFactory.create_backend("b200", device="cuda:0", dtype="float32")
eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
q = tb.QCTN(graph, backend=be)
for k, v in cores.items():
    q.cores_weights[k] = v.cuda().contiguous()
st = [s.cuda() for s in oc.unit_states(nq, K)]
torch.manual_seed(42)
x = torch.randn(B, nq, device="cuda")
for it in range(4):
    out = eng.contract_from_x(q, st, x, K=K)
torch.cuda.synchronize()
print("ok", float(out.sum()))
