import sys; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, tneq_b200
H = tneq_b200.QCTNHelper
for K, n in [(8,6),(16,6),(32,4),(32,6),(64,4)]:
    graph = H.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
    be = tneq_b200.BackendFactory.create_backend("b200", device="cuda:0", dtype="complex64")
    eng = tneq_b200.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    q = tneq_b200.QCTN(graph)
    torch.manual_seed(0)
    for c in q.cores:
        m = torch.randn(K*K, K*K, dtype=torch.complex128, device="cuda")
        qm, _ = torch.linalg.qr(m)
        q.cores_weights[c] = qm.reshape(K,K,K,K).to(torch.complex64)
    st = [torch.zeros(K, dtype=torch.complex64, device="cuda") for _ in range(n)]
    for s in st: s[-1] = 1.0
    B=3
    eye = torch.eye(K, dtype=torch.complex64, device="cuda").expand(B,K,K)
    import os
    os.environ["TNQ_FORCE_GEMM_PATH"]="1"
    got = eng.contract_with_compiled_strategy(q, st, [eye]*n)
    print(K, n, 'gemm path:', got.cpu().tolist())
