"""host-side cost of one training step through the public API vs device time (cfg3 network)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import __graft_entry__ as ge
ge.build()
import tneq_b200 as tb
dev = torch.device("cuda:0")
K, n = 3, 24
graph = bench.build_graph(tb, "merged", n, K)
for B in (2048, 16384):
    names, table, nq, cores_cpu, x = bench.synth_inputs(graph, K, B, "float32")
    be = tb.BackendFactory.create_backend("b200", device="cuda:0", dtype="float32")
    eng = tb.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    q = tb.QCTN(graph, backend=be)
    for c in names:
        q.cores_weights[c] = cores_cpu[c].to(dev).requires_grad_(True)
    states = [torch.zeros(K, device=dev) for _ in range(nq)]
    for s in states: s[-1] = 1.0
    mx, _ = eng.generate_data(x.to(dev), K=K, ret_type="TNTensor")
    mx = [tb.TNTensor(m.tensor.contiguous(), m.scale, m.log_scale) for m in mx]
    fn = eng._compiled(q, states, mx, True, "symmetric")
    cd = {c: q.cores_weights[c] for c in names}
    for _ in range(5): fn.loss_and_grads(cd, states, mx)
    torch.cuda.synchronize()
    N = 50
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N): fn.loss_and_grads(cd, states, mx)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B}: host issue {1e3*(t1-t0)/N:.3f} ms/step, device span {e0.elapsed_time(e1)/N:.3f} ms/step, wall incl sync {1e3*(t2-t0)/N:.3f}")
    for _ in range(5): eng.contract_with_compiled_strategy_for_gradient(q, states, mx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N): eng.contract_with_compiled_strategy_for_gradient(q, states, mx)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"   engine API: host issue {1e3*(t1-t0)/N:.3f} ms/step, wall incl sync {1e3*(t2-t0)/N:.3f}")
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20): fn.loss_and_grads(cd, states, mx)
    pr.disable(); torch.cuda.synchronize()
    if B == 2048: pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
