"""GPU parity of the warp-level ladder kernel (csrc/tnq_ladder.cu): two-layer merged MPS networks
(BASELINE cfg3) through the engine API and the C ABI, against the oracle at sizes the oracle
finishes in seconds, against the generic contraction VM (an independent CUDA implementation of the
same plan), and -- at cfg3's full size -- through size-independent properties.
Tolerance: 1e-5 relative (north_star) on well-conditioned samples, float64 oracle as yardstick.
"""
import os

import pytest
import torch

import tneq_b200
from oracle import qctn_oracle as oc
from helpers import make_case, well_conditioned_case, upcast, clone_mx, rel_err, elem_rel_err, NOISE_FACTOR

pytestmark = pytest.mark.gpu
H = tneq_b200.QCTNHelper
DEV = "cuda:0"


def merged_graph(n, K):
    q = tneq_b200.QCTN(H.generate_example_graph(n=n, graph_type="mps", dim_char=str(K)))
    return tneq_b200.QCTN.merge(q, q).graph


def _to_dev(x):
    if isinstance(x, oc.TNT):
        return tneq_b200.TNTensor(x.tensor.to(DEV), x.scale, x.log_scale)
    return x.to(DEV)


def _setup(graph, K, cores):
    be = tneq_b200.BackendFactory.create_backend("b200", device=DEV, dtype="float32")
    eng = tneq_b200.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    q = tneq_b200.QCTN(graph, backend=be)
    for k, v in cores.items():
        # contiguous: a CUDA graph is only captured over the caller's own buffers (no hidden copies)
        q.cores_weights[k] = v.to(DEV).contiguous().requires_grad_(True)
    return eng, q


def _bound(eng, q, st, mx):
    return next(iter(eng._compiled(q, st, mx, True, "symmetric").plans.values()))


@pytest.fixture(params=["warp-level", "second-generation"])
def generation(request):
    """Edge rank 3 has two implementations of the sweep: csrc/tnq_ladder.cu (default) and, with
    TNQ_LADDER_V2=1, csrc/tnq_ladder2.cu (lanes = samples, constant-memory operands, recomputed
    environment).  Both are held to the same parity bar."""
    if request.param == "second-generation":
        os.environ["TNQ_LADDER_V2"] = "1"
    yield request.param
    os.environ.pop("TNQ_LADDER_V2", None)


@pytest.mark.parametrize("n,K,B", [(3, 3, 4), (5, 3, 100), (24, 3, 50), (9, 2, 333), (24, 2, 64), (7, 3, 150)])
def test_ladder_vs_oracle(n, K, B, built_lib, generation):
    if generation == "second-generation" and K != 3:
        pytest.skip("the second-generation kernel covers edge rank 3")
    graph = merged_graph(n, K)
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=n + K)
    want = oc.forward(graph, cores, states, clone_mx(mxs))
    c64 = {k: v.double() for k, v in cores.items()}
    s64 = [s.double() for s in states]
    truth = oc.forward(graph, c64, s64, [upcast(m, torch.float64) for m in clone_mx(mxs)])
    eng, q = _setup(graph, K, cores)
    st = [s.to(DEV) for s in states]
    got = eng.contract_with_compiled_strategy(q, st, [_to_dev(m) for m in clone_mx(mxs)])
    assert _bound(eng, q, st, [_to_dev(m) for m in clone_mx(mxs)]).ladder is not None
    assert got.shape == want.shape and got.dtype == want.dtype
    assert elem_rel_err(got.double(), truth) < max(1e-5, NOISE_FACTOR * elem_rel_err(want.double(), truth))
    wl, wg = oc.loss_and_grads(graph, cores, states, clone_mx(mxs))
    tl, tg = oc.loss_and_grads(graph, c64, s64, [upcast(m, torch.float64) for m in clone_mx(mxs)])
    for fused in (True, False):   # fused loss kernel (mode 1), then the torch.autograd route (modes 0 + 2)
        loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, st, [_to_dev(m) for m in clone_mx(mxs)],
                                                                       fused=fused)
        assert abs(loss.item() - tl.item()) <= max(1e-5 * abs(tl.item()), NOISE_FACTOR * abs(wl.item() - tl.item()))
        assert len(grads) == len(wg)
        for g, w, t in zip(grads, wg, tg):
            assert g.shape == w.shape and g.dtype == w.dtype
            assert rel_err(g.double(), t) < max(1e-5, NOISE_FACTOR * rel_err(w.double(), t)), fused


def test_ladder_vs_vm_route(built_lib):
    """The generic contraction VM runs the same plan with different kernels and association order."""
    n, K, B = 24, 3, 1000
    graph = merged_graph(n, K)
    names, table, nq, cores, states, mxs = make_case(graph, K, B, "float32", seed=7)
    st = [s.to(DEV) for s in states]
    out = {}
    for route in ("ladder", "vm"):
        if route == "vm":
            os.environ["TNQ_NO_CHAIN"] = "1"
        try:
            eng, q = _setup(graph, K, cores)
            vals = eng.contract_with_compiled_strategy(q, st, [_to_dev(m) for m in clone_mx(mxs)])
            assert (_bound(eng, q, st, [_to_dev(m) for m in clone_mx(mxs)]).ladder is not None) == (route == "ladder")
            loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, st, [_to_dev(m) for m in clone_mx(mxs)])
            out[route] = (vals.cpu(), loss.item(), [g.cpu() for g in grads])
        finally:
            os.environ.pop("TNQ_NO_CHAIN", None)
    va, la, ga = out["ladder"]
    vb, lb, gb = out["vm"]
    big = vb.abs() > 1e-3 * vb.abs().max()           # samples that did not cancel in float32
    assert ((va - vb).abs()[big] / vb.abs()[big]).max() < 2e-4
    assert rel_err(va, vb) < 1e-5
    assert abs(la - lb) < 1e-4 * abs(lb)


def test_long_chain(built_lib, generation):
    """Long chains.  Random-data values underflow float32 at these lengths, so the checks are the
    normalisation known answer (orthogonal cores + identity measurements => 1), agreement of the two
    CUDA routes at 40 qubits (the VM's operand table ends at 192 inputs), and at 64 qubits -- where
    the constant pool no longer leaves room for the full warp count -- batch-split consistency."""
    K, B = 3, 100
    for n, against_vm in ((40, True), (64, False)):
        graph = merged_graph(n, K)
        torch.manual_seed(5)
        names, table, nq = oc.parse_graph(graph)
        cores = oc.random_cores(table, torch.float32)
        states = [s.to(DEV) for s in oc.unit_states(nq, K)]
        eye = [torch.eye(K, device=DEV).expand(B, K, K).contiguous() for _ in range(nq)]
        mx = [e + 0.05 * torch.randn(B, K, K, device=DEV) for e in eye]
        res = {}
        for route in ("ladder", "vm") if against_vm else ("ladder",):
            if route == "vm":
                os.environ["TNQ_NO_CHAIN"] = "1"
            try:
                eng, q = _setup(graph, K, cores)
                ones = eng.contract_with_compiled_strategy(q, states, eye)
                assert (ones - 1).abs().max().item() < 2e-5
                loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, states, mx)
                res[route] = (loss.item(), [g.cpu() for g in grads])
                if not against_vm:
                    la, ga = eng.contract_with_compiled_strategy_for_gradient(q, states, [m[:37] for m in mx])
                    lb, gb = eng.contract_with_compiled_strategy_for_gradient(q, states, [m[37:] for m in mx])
                    assert abs(loss.item() - (0.37 * la.item() + 0.63 * lb.item())) < 1e-5 * max(1.0, abs(loss.item()))
                    for f, u, v in zip(grads, ga, gb):
                        mix = 0.37 * u + 0.63 * v
                        assert torch.isfinite(f).all()
                        assert (f - mix).abs().max().item() <= 2e-4 * mix.abs().max().item() + 1e-7
            finally:
                os.environ.pop("TNQ_NO_CHAIN", None)
        if against_vm:
            assert abs(res["ladder"][0] - res["vm"][0]) < 1e-4 * max(1.0, abs(res["vm"][0]))
            for a, b in zip(res["ladder"][1], res["vm"][1]):
                assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item() + 1e-7


def test_cfg3_full_size_properties(built_lib, generation):
    """cfg3 size (24 qubits, batch 16384): normalisation known answer, ragged batches, and the
    loss / gradient of a batch equal to the sample-weighted mean over its two halves."""
    n, K, B = 24, 3, 16384
    graph = merged_graph(n, K)
    torch.manual_seed(11)
    names, table, nq = oc.parse_graph(graph)
    cores = oc.random_cores(table, torch.float32)
    states = [s.to(DEV) for s in oc.unit_states(nq, K)]
    eng, q = _setup(graph, K, cores)
    # KAT-1: orthogonal cores + identity measurements => 1 for every sample
    eye = [torch.eye(K, device=DEV).expand(B, K, K).contiguous() for _ in range(nq)]
    got = eng.contract_with_compiled_strategy(q, states, eye)
    assert got.shape == (B,)
    assert (got - 1).abs().max().item() < 1e-5
    # random data: the batch against its two (ragged) halves
    x = torch.randn(B, nq, device=DEV)
    mx, _ = eng.generate_data(x, K=K, ret_type="tensor")
    mx = [m.contiguous() * 3.0 for m in mx]           # keep typical values away from the 1e-10 clamp
    full = eng.contract_with_compiled_strategy(q, states, mx)
    # 64 random samples of THIS batch against the oracle (the reference's algorithm on CPU, float32
    # and float64): the full-size launch computes the same per-sample numbers as the small ones
    gen = torch.Generator().manual_seed(2)
    pick = torch.randperm(B, generator=gen)[:64].sort().values
    sub = [m[pick.to(DEV)].cpu() for m in mx]
    st_cpu = oc.unit_states(nq, K)
    want = oc.forward(graph, cores, st_cpu, [m.clone() for m in sub])
    truth = oc.forward(graph, {k: v.double() for k, v in cores.items()}, [s.double() for s in st_cpu],
                       [m.double() for m in sub])
    got = full[pick.to(DEV)].cpu()
    assert rel_err(got, want) < 1e-5
    assert rel_err(got.double(), truth) < max(1e-5, NOISE_FACTOR * rel_err(want.double(), truth))
    cut = 7001
    a = eng.contract_with_compiled_strategy(q, states, [m[:cut] for m in mx])
    b = eng.contract_with_compiled_strategy(q, states, [m[cut:] for m in mx])
    assert torch.equal(full, torch.cat([a, b]))       # per-sample arithmetic does not depend on the batch
    lf, gf = eng.contract_with_compiled_strategy_for_gradient(q, states, mx)
    la, ga = eng.contract_with_compiled_strategy_for_gradient(q, states, [m[:cut] for m in mx])
    lb, gb = eng.contract_with_compiled_strategy_for_gradient(q, states, [m[cut:] for m in mx])
    wa, wb = cut / B, (B - cut) / B
    assert abs(lf.item() - (wa * la.item() + wb * lb.item())) < 1e-5 * abs(lf.item())
    for f, u, v in zip(gf, ga, gb):
        mix = wa * u + wb * v
        assert (f - mix).abs().max().item() <= 2e-4 * mix.abs().max().item() + 1e-7
    # determinism: the same launch twice gives bit-identical gradients
    lf2, gf2 = eng.contract_with_compiled_strategy_for_gradient(q, states, mx)
    assert lf.item() == lf2.item() and all(torch.equal(u, v) for u, v in zip(gf, gf2))


def test_cuda_graph_replay(built_lib):
    """Opt-in CUDA-graph replay of the fused training step: same numbers as the direct launches,
    follows in-place updates of the cores and new data copied into the same input buffers, and
    falls back to direct launches when the buffers change."""
    n, K, B = 24, 3, 500
    graph = merged_graph(n, K)
    names, table, nq, cores, states, mxs = make_case(graph, K, B, "float32", seed=3)
    eng, q = _setup(graph, K, cores)
    st = [s.to(DEV) for s in states]
    mx = [_to_dev(m) for m in clone_mx(mxs)]
    l0, g0 = eng.contract_with_compiled_strategy_for_gradient(q, st, mx)
    l0, g0 = l0.item(), [g.clone() for g in g0]
    fn = eng._compiled(q, st, mx, True, "symmetric")
    eng.enable_cuda_graphs(True)
    try:
        for it in range(4):                        # 1: direct, 2: capture, 3+: replay
            l, g = eng.contract_with_compiled_strategy_for_gradient(q, st, mx)
            assert l.item() == pytest.approx(l0, rel=1e-6)
            assert all(torch.equal(a, b) for a, b in zip(g, g0))
        assert fn.graph_stats["replays"] >= 2
        # in-place core update and new measurement data in the same buffers
        with torch.no_grad():
            q.cores_weights[names[3]].mul_(1.01)
            for m in mx:
                m.tensor.mul_(0.97)
        lr, gr = eng.contract_with_compiled_strategy_for_gradient(q, st, mx)
        lr, gr = lr.item(), [g.clone() for g in gr]
        eng.enable_cuda_graphs(False)
        ld, gd = eng.contract_with_compiled_strategy_for_gradient(q, st, mx)
        assert lr == pytest.approx(ld.item(), rel=1e-6)
        assert all(torch.equal(a, b) for a, b in zip(gr, gd))
        # different buffers: no stale replay
        eng.enable_cuda_graphs(True)
        mx2 = [tneq_b200.TNTensor(m.tensor.clone() * 1.1, m.scale, m.log_scale) for m in mx]
        l2, g2 = eng.contract_with_compiled_strategy_for_gradient(q, st, mx2)
        eng.enable_cuda_graphs(False)
        l3, g3 = eng.contract_with_compiled_strategy_for_gradient(q, st, mx2)
        assert l2.item() == pytest.approx(l3.item(), rel=1e-6)
        assert all(torch.equal(a, b) for a, b in zip(g2, g3))
    finally:
        eng.enable_cuda_graphs(False)


def test_forward_graph_replay_chain_and_ladder(built_lib):
    """Opt-in CUDA-graph replay of the FORWARD launches (BASELINE cfg2: 16-qubit MPS, batch 4096, is a
    10 us kernel behind ~50 us of per-call host work): same values as the direct launch, follows data
    copied into the same buffers, no stale replay when the buffers change."""
    for kind, n, K, B in (("mps", 16, 3, 4096), ("merged", 8, 3, 300)):
        graph = merged_graph(n, K) if kind == "merged" else H.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
        names, table, nq, cores, states, mxs = make_case(graph, K, B, "float32", seed=4)
        eng, q = _setup(graph, K, cores)
        st = [s.to(DEV) for s in states]
        mx = [_to_dev(m) for m in clone_mx(mxs)]
        fn = eng._compiled(q, st, mx, True, "symmetric")
        cd = {c: q.cores_weights[c] for c in names}
        with torch.no_grad():
            want = fn(cd, st, mx).tensor.clone()
            eng.enable_cuda_graphs(True)
            try:
                for it in range(4):                    # 1: direct (first sighting), 2: capture, 3+: replay
                    got = fn(cd, st, mx)
                    assert torch.equal(got.tensor, want)
                assert fn.graph_stats["replays"] >= 2
                for m in mx:
                    m.tensor.mul_(0.9)
                rep = fn(cd, st, mx).tensor.clone()
                eng.enable_cuda_graphs(False)
                direct = fn(cd, st, mx).tensor
                assert torch.equal(rep, direct) and not torch.equal(rep, want)
                eng.enable_cuda_graphs(True)
                mx2 = [tneq_b200.TNTensor(m.tensor.clone(), m.scale, m.log_scale) for m in mx]
                other = fn(cd, st, mx2).tensor
                assert torch.equal(other, direct)
            finally:
                eng.enable_cuda_graphs(False)
