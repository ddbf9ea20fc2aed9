"""Host-side mirror of the reference API: data model, registries, checkpoint format, C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import tneq_b200
from tneq_b200 import QCTN, QCTNHelper, TNTensor, BackendFactory, StrategyCompiler
from oracle import qctn_oracle as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_example_graphs_and_adjacency_match_the_oracle_parser():
    assert QCTNHelper.generate_example_graph(n=3, graph_type="mps", dim_char="2") == \
        "-2-a-------2-\n-2-a--2--b-2-\n-2-------b-2-\n"
    assert QCTNHelper.generate_example_graph(n=4, graph_type="tree", dim_char="3") == \
        "-3-----a-3-\n-3-b-3-a-3-\n-3-b-3-c-3-\n-3-----c-3-\n"
    for kind in ("mps", "tree", "wall"):
        for n in (4, 5, 8):
            g = QCTNHelper.generate_example_graph(n=n, graph_type=kind, dim_char="2")
            q = QCTN(g)
            names, table, nq = oc.parse_graph(g)
            assert q.cores == names and q.nqubits == nq
            for t, o in zip(q.adjacency_table, table):
                conv = lambda es: [(e["neighbor_idx"], e["edge_rank"], e["qubit_idx"]) for e in es]
                assert conv(t["in_edge_list"]) == [(e["nbr"], e["rank"], e["qubit"]) for e in o["ins"]]
                assert conv(t["out_edge_list"]) == [(e["nbr"], e["rank"], e["qubit"]) for e in o["outs"]]


def test_merge_and_split_round_trip():
    q = QCTN(QCTNHelper.generate_example_graph(n=4, graph_type="mps", dim_char="2"))
    m = QCTN.merge(q, q)
    assert m.ncores == 6 and m.nqubits == 4
    assert m.graph.splitlines()[1] == "-2-a--2--b-------2-d--2--e-------2-"
    left, right = m.split()
    assert left.cores == ["a", "b", "c"] and right.cores == ["d", "e", "f"]
    assert [t["input_shape"] + t["output_shape"] for t in left.adjacency_table] == [[2, 2, 2, 2]] * 3
    with pytest.raises(ValueError):
        m.split(0)


class _CpuBackend:
    """just enough of a backend for the data-model tests (no contraction)"""
    def init_random_core(self, shape):
        return oc.init_random_core(shape)
    def reshape(self, t, shape):
        return t.reshape(shape)
    def tensor_to_numpy(self, t):
        return t.detach().numpy()
    def convert_to_tensor(self, a):
        return torch.as_tensor(a)


def test_set_cores_and_checkpoint_format(tmp_path):
    be = _CpuBackend()
    q = QCTN(QCTNHelper.generate_example_graph(n=4, graph_type="mps", dim_char="2"), backend=be)
    assert all(tuple(q.cores_weights[c].shape) == (2, 2, 2, 2) for c in q.cores)
    new = [torch.randn(4, 4) for _ in q.cores]
    q.set_cores(new)
    assert tuple(q.cores_weights["a"].shape) == (2, 2, 2, 2)
    with pytest.raises(ValueError):
        q.set_cores(new[:2])
    with pytest.raises(ValueError):
        q.set_cores({"a": new[0]})
    with pytest.raises(ValueError):
        q.set_cores([torch.randn(3, 3)] * 3)
    with pytest.raises(TypeError):
        q.set_cores(42)
    path = str(tmp_path / "cores.safetensors")
    q.cores_weights["b"] = TNTensor(q.cores_weights["b"] / 4.0, 4.0)
    q.save_cores(path, metadata={"step": 7})
    from safetensors.numpy import load_file
    blob = load_file(path)
    assert sorted(blob) == ["core_a", "core_b", "core_c"]          # reference key layout (qctn.py:902-926)
    q2 = QCTN.from_pretrained(q.graph, path, backend=be)
    for c in q.cores:
        w = q2.cores_weights[c]
        assert isinstance(w, TNTensor)                              # loaded cores are auto-scaled TNTensors
        ref = q.cores_weights[c]
        ref = ref.tensor * ref.scale if isinstance(ref, TNTensor) else ref
        assert torch.allclose(w.tensor * w.scale, ref, atol=1e-6)
    assert q2._loaded_metadata == {"step": "7"}


def test_tntensor_scaling_keeps_the_value():
    t = TNTensor(torch.tensor([0.5, -2.0, 1.0]))
    t.auto_scale()
    assert t.scale == 2.0 and torch.allclose(t.tensor, torch.tensor([0.25, -1.0, 0.5]))
    t.scale_with(4.0)
    assert torch.allclose(t.tensor * t.scale, torch.tensor([0.5, -2.0, 1.0]))
    t.scale_to(1.0)
    assert t.scale == 1.0 and t.log_scale == 0.0 and torch.allclose(t.tensor, torch.tensor([0.5, -2.0, 1.0]))
    with pytest.raises(ValueError):
        t.scale_to(0)
    assert not t.is_complex() and t.conj().scale == 1.0


def test_registries_and_no_cpu_fallback():
    assert "b200" in BackendFactory._backends
    for mode in ("fast", "balanced", "full"):
        assert "b200" in StrategyCompiler.MODES[mode]
    with pytest.raises(ValueError):
        BackendFactory.create_backend("jax")
    with pytest.raises(ValueError):
        StrategyCompiler(mode="bogus")
    strat = StrategyCompiler.get_registered_strategies()["b200"]
    assert strat.estimate_cost(None, {}) < 5e5                      # beats GreedyStrategy's 5e5 (greedy_strategy.py:608)
    with pytest.raises(RuntimeError):
        BackendFactory.create_backend("b200", device="cpu")        # the product never computes on the CPU
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            BackendFactory.create_backend("b200", device="cuda")


def test_compute_function_refuses_cpu_tensors():
    q = QCTN(QCTNHelper.generate_example_graph(n=3, graph_type="mps", dim_char="2"), backend=_CpuBackend())
    strat = StrategyCompiler.get_registered_strategies()["b200"]
    fn = strat.get_compute_function(q, {}, None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fn({c: q.cores_weights[c] for c in q.cores}, oc.unit_states(3, 2), [torch.randn(2, 2, 2)] * 3)
    with pytest.raises(ValueError):
        strat.get_compute_function(q, {}, None, right_qctn=3.14)
    assert fn.equations(oc.unit_states(3, 2), [torch.randn(2, 2, 2)] * 3)[-1] == "acd,adc->a"


def test_c_abi_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "tneq_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tnq_[a-z0-9_]+)\s*\(", header)))
    assert "tnq_plan_run" in declared and "tnq_plan_create" in declared
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/tneq_b200.h but not exported"
    from tneq_b200 import _lib
    assert sorted(_lib.EXPORTS) == [d for d in declared if d in _lib.EXPORTS]
    lib.tnq_last_error.restype = ctypes.c_char_p
    # argument validation happens before any CUDA call: safe without a GPU
    assert lib.tnq_plan_create(None, 0, None) != 0
    assert b"bad arguments" in lib.tnq_last_error()
    bad = (ctypes.c_int64 * 16)(*([1] * 16))
    out = ctypes.c_void_p()
    assert lib.tnq_plan_create(bad, 16, ctypes.byref(out)) != 0
    assert b"magic" in lib.tnq_last_error()


def test_optimizer_steps_match_the_oracle_on_cpu():
    """SGDG / Cayley step (backend_pytorch.py:349-468) -- device-agnostic torch code."""
    import random
    from tneq_b200.optim import steps
    torch.manual_seed(0)
    for dtype in (torch.float64, torch.complex128):
        params = [oc.init_random_core([4, 4], dtype).reshape(2, 2, 2, 2) for _ in range(3)]
        grads = [torch.randn(2, 2, 2, 2, dtype=dtype) for _ in range(3)]
        random.seed(5)
        want, st_w = oc.sgdg_step([p.clone() for p in params], grads, {}, lr=0.05, momentum=0.9)
        random.seed(5)
        got, st_g = steps.optimizer_update([p.clone() for p in params], grads, {}, "sgdg",
                                           dict(learning_rate=0.05, momentum=0.9, stiefel=True))
        for a, b in zip(got, want):
            assert torch.allclose(a, b, atol=1e-12)
        for a, b in zip(st_g["momentum_buffer"], st_w["momentum_buffer"]):
            assert torch.allclose(a, b, atol=1e-12)
        m = got[0].reshape(4, 4)
        assert torch.allclose(m @ m.conj().T, torch.eye(4, dtype=dtype), atol=1e-6)     # stays on the manifold
    with pytest.raises(ValueError):
        steps.optimizer_update([], [], {}, "lbfgs", {})


def test_reference_plugin_registration():
    """With the reference importable (build container only), our classes plug into ITS registries
    and its StrategyCompiler selects 'b200' over 'greedy'."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present on this machine")
    ns = rh.load()
    import tneq_b200.reference_plugin as plug
    RefBackend, RefStrategy = plug.register()
    assert issubclass(RefBackend, ns.ComputeBackend) and issubclass(RefStrategy, ns.ContractionStrategy)
    assert "b200" in ns.BackendFactory._backends
    assert "b200" in ns.StrategyCompiler.MODES["balanced"] and "greedy" in ns.StrategyCompiler.MODES["balanced"]
    be, _ = rh.make_engine()
    with rh.quiet():
        q = ns.QCTN(ns.QCTNHelper.generate_example_graph(n=4, graph_type="mps", dim_char="2"), backend=be)
        # a network that lives on the reference's CPU backend stays with the reference's strategies
        fn, name, cost = ns.StrategyCompiler(mode="balanced").compile(q, {}, be, right_qctn="symmetric")
        assert name == "greedy"

        class OnB200:                      # (B200Backend itself cannot be constructed without a GPU)
            def get_backend_name(self):
                return "b200"
        q.backend = OnB200()
        fn, name, cost = ns.StrategyCompiler(mode="balanced").compile(q, {}, be, right_qctn="symmetric")
    assert name == "b200" and cost < 5e5
    assert fn.equations(oc.unit_states(4, 2), [torch.randn(2, 2, 2)] * 4)[0] == "cdef,c,aeg,higj,h,d,i->ajf"
    # leave the reference's registry as we found it for the other tests
    ns.StrategyCompiler.MODES["balanced"].remove("b200")
    ns.StrategyCompiler.MODES["full"].remove("b200")


def test_core_only_expression_matches_reference():
    """EngineSiamese.build_core_only_expression reproduces the reference's einsum bookkeeping for the cores-only
    contraction (einsum_strategy.py:137-194) on MPS, tree, wall and merged graphs (reference checkout needed)."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present on this machine")
    ns = rh.load()
    import importlib
    es = importlib.import_module("tneq_qc.contractor.einsum_strategy")
    builder = next(getattr(es, n) for n in dir(es) if hasattr(getattr(es, n), "build_core_only_expression"))
    be, _ = rh.make_engine()
    for kind, n, K in [("mps", 5, 2), ("tree", 6, 2), ("wall", 4, 2), ("mps", 4, 3)]:
        graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type=kind, dim_char=str(K))
        with rh.quiet():
            qr = ns.QCTN(graph, backend=be)
        want_eq, want_shapes = builder.build_core_only_expression(qr)
        q = tneq_b200.QCTN(graph)
        for c in q.cores:
            q.cores_weights[c] = torch.zeros(q.core_shape(c))
        got_eq, got_shapes = tneq_b200.EngineSiamese.build_core_only_expression(q)
        assert got_eq == want_eq, (kind, got_eq, want_eq)
        assert [tuple(s_) for s_ in got_shapes] == [tuple(s_) for s_ in want_shapes]
