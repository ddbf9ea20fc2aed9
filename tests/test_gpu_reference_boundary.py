"""The drop-in boundary exercised for real: the REFERENCE's own EngineSiamese / QCTN / Optimizer
(an unmodified copy of /root/reference/tneq_qc staged by build() under the git-ignored
baseline/_ref/, imported through oracle/ref_harness.py) drive this package's CUDA path after
`tneq_b200.reference_plugin.register()`, and the results are compared, in the same process, with
the reference's own 'pytorch' CPU backend + GreedyStrategy on the same inputs.

Covers backend_factory.py:91-100 (register_backend), compiler.py:38-54,110,123 (register_strategy,
kwarg forwarding, min-cost selection), engine_siamese.py:261-349 (forward), :351-554 (loss and
gradients through the reference's backend.compute_value_and_grad), :584-645 (marginal), :740-915
(sample), and `right_qctn=<QCTN>` (engine_siamese.py:304,390).
"""
import contextlib
import io

import pytest
import torch

from oracle import qctn_oracle as oc
from helpers import well_conditioned_case, clone_mx, rel_err, NOISE_FACTOR

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref(built_lib):
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("no copy of the reference package on this machine (baseline/_ref is staged by build())")
    ns = rh.load()
    import tneq_b200.reference_plugin as plug
    plug.register()
    yield rh, ns
    for mode in ("balanced", "full"):
        if "b200" in ns.StrategyCompiler.MODES[mode]:
            ns.StrategyCompiler.MODES[mode].remove("b200")


def _engines(ns, K, dtype="float32"):
    with contextlib.redirect_stdout(io.StringIO()):
        be_cpu = ns.BackendFactory.create_backend("pytorch", device="cpu", dtype=dtype)
        be_gpu = ns.BackendFactory.create_backend("b200", device=DEV, dtype=dtype)
        eng_cpu = ns.EngineSiamese(backend=be_cpu, strategy_mode="balanced", mx_K=K)
        eng_gpu = ns.EngineSiamese(backend=be_gpu, strategy_mode="balanced", mx_K=K)
    return be_cpu, eng_cpu, be_gpu, eng_gpu


def _networks(ns, graph, cores, be_cpu, be_gpu, grad=True):
    with contextlib.redirect_stdout(io.StringIO()):
        qc, qg = ns.QCTN(graph, backend=be_cpu), ns.QCTN(graph, backend=be_gpu)
    for k, v in cores.items():
        qc.cores_weights[k] = v.clone().requires_grad_(grad)
        qg.cores_weights[k] = v.to(DEV).requires_grad_(grad)
    return qc, qg


def _strategy_of(q):
    names = [getattr(q, a)["strategy_name"] for a in dir(q) if a.startswith("_compiled_strategy_")]
    assert names
    return set(names)


def _mx(ns, mxs, dev):
    out = []
    for m in clone_mx(mxs):
        if isinstance(m, oc.TNT):
            out.append(ns.TNTensor(m.tensor.to(dev), m.scale, m.log_scale))
        else:
            out.append(m.to(dev))
    return out


@pytest.mark.parametrize("kind,n,K,B,dtype", [("mps", 6, 3, 40, "float32"), ("merged", 5, 3, 24, "float32"),
                                             ("tree", 6, 2, 30, "float32"), ("mps", 5, 3, 12, "complex64")])
def test_reference_engine_on_b200_backend(ref, kind, n, K, B, dtype):
    rh, ns = ref
    import tneq_b200
    H = tneq_b200.QCTNHelper
    if kind == "merged":
        g1 = tneq_b200.QCTN(H.generate_example_graph(n=n, graph_type="mps", dim_char=str(K)))
        graph = tneq_b200.QCTN.merge(g1, g1).graph
    else:
        graph = H.generate_example_graph(n=n, graph_type=kind, dim_char=str(K))
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, dtype, seed=n)
    be_cpu, eng_cpu, be_gpu, eng_gpu = _engines(ns, K, dtype)
    qc, qg = _networks(ns, graph, cores, be_cpu, be_gpu)
    st_gpu = [s.to(DEV) for s in states]
    launches0 = tneq_b200._lib.launch_count()
    with contextlib.redirect_stdout(io.StringIO()):
        want = eng_cpu.contract_with_compiled_strategy(qc, states, _mx(ns, mxs, "cpu"))
        got = eng_gpu.contract_with_compiled_strategy(qg, st_gpu, _mx(ns, mxs, DEV))
    assert _strategy_of(qc) == {"greedy"} and _strategy_of(qg) == {"b200"}
    assert tneq_b200._lib.launch_count() > launches0, "the CUDA library did not run"
    assert got.device.type == "cuda" and got.shape == want.shape
    tol = 2e-5 if "complex" in dtype else 1e-5
    assert rel_err(got, want) < tol
    with contextlib.redirect_stdout(io.StringIO()):
        wl, wg = eng_cpu.contract_with_compiled_strategy_for_gradient(qc, states, _mx(ns, mxs, "cpu"))
        gl, gg = eng_gpu.contract_with_compiled_strategy_for_gradient(qg, st_gpu, _mx(ns, mxs, DEV))
    assert abs(gl.item() - wl.item()) <= 1e-5 * abs(wl.item())
    assert len(gg) == len(wg)
    # float64 yardstick (oracle == reference bit for bit, tests/golden): as close to it as the
    # reference's own float32 arithmetic is
    td64 = torch.complex128 if "complex" in dtype else torch.float64
    from helpers import upcast
    tl, tg = oc.loss_and_grads(graph, {k: v.to(td64) for k, v in cores.items()}, [s.to(td64) for s in states],
                               [upcast(m, td64) for m in clone_mx(mxs)])
    for g, w, t in zip(gg, wg, tg):
        assert g.shape == w.shape and g.dtype == w.dtype and g.device.type == "cuda"
        ref_err = rel_err(w.to(td64), t)
        assert rel_err(g.to(td64), t) < max(1e-5, NOISE_FACTOR * ref_err), (rel_err(g.to(td64), t), ref_err)


def test_reference_optimizer_step_and_probabilities(ref):
    """Optimizer.step of the reference on the b200 backend (tnq_sgdg_step underneath), then the
    reference's marginal / conditional identity (tests/test_probabilities.py:84-87) and sample()
    (:296-333: shape and bounds) on the device."""
    rh, ns = ref
    import random
    import tneq_b200
    K, n, B = 3, 4, 16
    graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=1)
    be_cpu, eng_cpu, be_gpu, eng_gpu = _engines(ns, K)
    qc, qg = _networks(ns, graph, cores, be_cpu, be_gpu)
    st_gpu = [s.to(DEV) for s in states]
    with contextlib.redirect_stdout(io.StringIO()):
        opt_c = ns.Optimizer(method="sgdg", learning_rate=0.05, engine=eng_cpu, momentum=0.9, stiefel=True)
        opt_g = ns.Optimizer(method="sgdg", learning_rate=0.05, engine=eng_gpu, momentum=0.9, stiefel=True)
        for it in range(3):
            lc, gc = eng_cpu.contract_with_compiled_strategy_for_gradient(qc, states, _mx(ns, mxs, "cpu"))
            lg, gg = eng_gpu.contract_with_compiled_strategy_for_gradient(qg, st_gpu, _mx(ns, mxs, DEV))
            assert abs(lg.item() - lc.item()) <= 2e-5 * abs(lc.item()), it
            random.seed(50 + it)
            opt_c.step(qc, gc)
            random.seed(50 + it)
            opt_g.step(qg, gg)
    for k in names:
        a, b = qg.cores_weights[k], qc.cores_weights[k]
        a = a.tensor * a.scale if hasattr(a, "scale") else a
        b = b.tensor * b.scale if hasattr(b, "scale") else b
        assert rel_err(a.detach(), b.detach()) < 2e-4, k
    # probabilities on the device, through the reference's own engine methods
    raw = [oc._raw(m).to(DEV) for m in clone_mx(mxs)]
    with contextlib.redirect_stdout(io.StringIO()):
        joint = eng_gpu.calculate_marginal_probability(qg, st_gpu, [raw[0], raw[1]], [0, 1])
        marg = eng_gpu.calculate_marginal_probability(qg, st_gpu, [raw[0]], [0])
        jc = eng_cpu.calculate_marginal_probability(qc, states, [raw[0].cpu(), raw[1].cpu()], [0, 1])
        both = torch.stack([raw[1], torch.eye(K, device=DEV).expand(B, K, K)], dim=1)
        ab = eng_gpu.contract_with_compiled_strategy(
            qg, st_gpu, [raw[0], both] + [torch.eye(K, device=DEV).expand(B, K, K)] * (n - 2))
    assert rel_err(joint, jc) < 1e-4       # (cores moved by three optimizer steps on both sides)
    assert ab.shape == (B, 2)
    assert torch.allclose(ab[:, 0] / (ab[:, 1] + 1e-10), joint / (marg + 1e-10), atol=1e-5)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        samples = eng_gpu.sample(qg, st_gpu, num_samples=40, K=K, bounds=[-5, 5], grid_size=30)
    assert tuple(samples.shape) == (40, n) and samples.device.type == "cuda"
    assert (samples >= -5).all() and (samples <= 5).all()


def test_right_qctn_given_as_second_network(ref):
    """`right_qctn=<QCTN>` (engine_siamese.py:304,390; greedy_strategy.py right_qctn branch): the
    reference's engine on both backends, forward values and the gradients of BOTH networks."""
    rh, ns = ref
    import tneq_b200
    K, n, B = 2, 4, 12
    graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=2)
    torch.manual_seed(77)
    rcores = {k: v + 0.1 * torch.randn_like(v) for k, v in cores.items()}
    be_cpu, eng_cpu, be_gpu, eng_gpu = _engines(ns, K)
    qc, qg = _networks(ns, graph, cores, be_cpu, be_gpu)
    rc, rg = _networks(ns, graph, rcores, be_cpu, be_gpu)
    st_gpu = [s.to(DEV) for s in states]
    raw = [oc._raw(m) for m in clone_mx(mxs)]
    with contextlib.redirect_stdout(io.StringIO()):
        want = eng_cpu.contract_with_compiled_strategy(qc, states, raw, right_qctn=rc)
        got = eng_gpu.contract_with_compiled_strategy(qg, st_gpu, [m.to(DEV) for m in raw], right_qctn=rg)
    assert rel_err(got, want) < 1e-5
    with contextlib.redirect_stdout(io.StringIO()):
        wl, wg = eng_cpu.contract_with_compiled_strategy_for_gradient(qc, states, raw, right_qctn=rc)
        gl, gg = eng_gpu.contract_with_compiled_strategy_for_gradient(qg, st_gpu, [m.to(DEV) for m in raw], right_qctn=rg)
    assert abs(gl.item() - wl.item()) <= 1e-5 * abs(wl.item())
    assert len(gg) == len(wg) == 2 * len(names)
    for g, w in zip(gg, wg):
        assert rel_err(g, w) < 5e-5


def test_checkpoint_interchange_and_core_only_contraction(ref, tmp_path):
    """the safetensors checkpoint layout (qctn.py:902-964) written by the reference loads here and
    vice versa, and the cores-only contraction (einsum_strategy.py:137-194 + engine.py:228-252) of the loaded
    network equals a float64 einsum of the same equation on the CPU."""
    rh, ns = ref
    import tneq_b200
    K, n = 2, 5
    graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="tree", dim_char=str(K))
    names, table, nq = oc.parse_graph(graph)
    torch.manual_seed(8)
    cores = oc.random_cores(table)
    be_cpu, eng_cpu, be_gpu, eng_gpu = _engines(ns, K)
    with contextlib.redirect_stdout(io.StringIO()):
        qref = ns.QCTN(graph, backend=be_cpu)
    for k, v in cores.items():
        qref.cores_weights[k] = v.clone()
    f1, f2 = str(tmp_path / "ref.safetensors"), str(tmp_path / "ours.safetensors")
    qref.save_cores(f1, metadata={"note": "written by the reference"})
    ours_be = tneq_b200.BackendFactory.create_backend("b200", device=DEV, dtype="float32")
    qo = tneq_b200.QCTN.from_pretrained(graph, f1, backend=ours_be)
    assert qo._loaded_metadata.get("note") == "written by the reference"
    for k, v in cores.items():
        w = qo.cores_weights[k]
        assert rel_err((w.tensor * w.scale).cpu(), v) < 1e-6
    qo.save_cores(f2)
    with contextlib.redirect_stdout(io.StringIO()):
        qback = ns.QCTN.from_pretrained(graph, f2, backend=be_cpu)
    for k, v in cores.items():
        w = qback.cores_weights[k]
        assert rel_err(w.tensor * w.scale, v) < 1e-6
    eng = tneq_b200.EngineSiamese(backend=ours_be, strategy_mode="balanced", mx_K=K)
    eq, shapes = eng.build_core_only_expression(qo)
    dense = eng.contract_core_only(qo)
    want = torch.einsum(eq, *[cores[k].double() for k in names])
    assert dense.device.type == "cuda" and tuple(dense.shape) == tuple(want.shape)
    assert rel_err(dense.double().cpu(), want) < 1e-5
