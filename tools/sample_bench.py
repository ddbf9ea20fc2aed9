"""EngineSiamese.sample on the 16-qubit MPS network (K = 3): the reference's procedure (method='grid': one forward per
qubit at batch S x G), method='linear' (K^2 contractions per sample and qubit) and method='prefix' (one kernel, prefix
environments).  python tools/sample_bench.py [S] [G]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tneq_b200
from oracle import qctn_oracle as oc

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n, K = 16, 3
graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
be = tneq_b200.BackendFactory.create_backend("b200", device="cuda:0", dtype="float32")
eng = tneq_b200.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
names, table, nq = oc.parse_graph(graph)
torch.manual_seed(1234)
cores = oc.random_cores(table)
q = tneq_b200.QCTN(graph, backend=be)
for k, v in cores.items():
    q.cores_weights[k] = v.cuda()
st = [s.cuda() for s in oc.unit_states(nq, K)]
out = {}
for method in ("prefix", "linear", "grid"):
    if method == "grid" and S * G > 4_000_000:
        continue
    for rep in range(2):
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out[method] = eng.sample(q, st, num_samples=S, K=K, bounds=[-5, 5], grid_size=G, method=method)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"sample(method={method!r}): S={S} G={G} n={n}: {dt * 1e3:.2f} ms ({S / dt:.0f} samples/s)", flush=True)
for m in out:
    oob = ((out[m] < -5) | (out[m] > 5)).sum().item()
    print(f"{m}: {oob} of {out[m].numel()} values outside the bounds (the reference's (u - c0) / (c1 - c0 + 1e-10) where the "
          f"float32 cdf has saturated: reference behaviour, all methods)")
for m in out:
    if m != "prefix":
        d = (out[m] - out["prefix"]).abs()
        print(f"prefix vs {m}: median |diff| {d.median().item():.2e}, max {d.max().item():.2e}")
