"""The ladder kernel's arithmetic (csrc/tnq_ladder_core.cuh) checked on the CPU.

tests/emu/ladder_emu.cpp compiles the SAME phase functions and sweep driver that the
CUDA kernel uses (csrc/tnq_ladder.cu) with the 32 lanes of a warp executed one after
the other, so the index bookkeeping of phases A/B/C, the reverse sweep, the per-warp
checkpoints and the lane reduction are validated against the oracle without a GPU.
This is test infrastructure: nothing in the product can reach the emulation.
"""
import ctypes
import math
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import tneq_b200
from tneq_b200.contractor.plan import ContractionPlan
from oracle import qctn_oracle as oc
from helpers import well_conditioned_case, clone_mx
from test_oracle_golden import load_case, fresh_mx, GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "quantum_circuits_symmetry_breaking_based_on_tneq-qc_b200", "csrc")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "ladder_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                    os.path.join(ROOT, "tests", "emu", "ladder_emu.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def merged_graph(n, K):
    q = tneq_b200.QCTN(tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K)))
    return tneq_b200.QCTN.merge(q, q).graph


def ladder_of(graph, K):
    q = tneq_b200.QCTN(graph)
    n = q.nqubits
    plan = ContractionPlan(q.adjacency_table, n, {c: (K,) * 4 for c in q.cores}, {i: K for i in range(n)},
                           {i: ("a", K, K) for i in range(n)}, "float32")
    lad = plan.mps_ladder()
    assert lad is not None and lad[0] == K
    return lad


def run_emu(lib, K, n, cores, layer1, layer2, states, ms, B, mode, seed=None, log_scale=0.0, nwarps=3, strides=None):
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    ca = [cores[k].contiguous() for k in layer1]
    cx = [cores[k].contiguous() for k in layer2]
    ga, gx = [torch.zeros_like(c) for c in ca], [torch.zeros_like(c) for c in cx]
    vals, loss = torch.zeros(B), torch.zeros(1)
    sd = seed if seed is not None else torch.zeros(B)
    st = (ctypes.c_longlong * n)(*(strides or [K * K] * n))
    rc = lib.ladder_emu(K, n, arr(ca), arr(cx), arr(states), arr(ms), st, ctypes.c_longlong(B), mode,
                        ctypes.c_void_p(sd.data_ptr()), ctypes.c_void_p(vals.data_ptr()),
                        ctypes.c_void_p(loss.data_ptr()), arr(ga), arr(gx), ctypes.c_double(log_scale), nwarps)
    assert rc == 0
    grads = dict(zip(layer1, ga))
    grads.update(zip(layer2, gx))
    return vals, loss[0], grads


def test_recogniser_rejects_other_networks():
    K = 3
    for kind in ("mps", "tree"):
        g = tneq_b200.QCTNHelper.generate_example_graph(n=6, graph_type=kind, dim_char=str(K))
        q = tneq_b200.QCTN(g)
        shapes = {c: tuple(q.cores_weights[c].shape) if hasattr(q, "cores_weights") and c in q.cores_weights else None
                  for c in q.cores}
        if any(v is None for v in shapes.values()):
            names, table, nq = oc.parse_graph(g)
            shapes = oc.core_shapes(table)
        plan = ContractionPlan(q.adjacency_table, q.nqubits, shapes, {i: K for i in range(6)},
                               {i: ("a", K, K) for i in range(6)}, "float32")
        assert plan.mps_ladder() is None
    # a missing measurement changes the greedy groups: not the ladder pattern any more
    q = tneq_b200.QCTN(merged_graph(6, K))
    plan = ContractionPlan(q.adjacency_table, 6, {c: (K,) * 4 for c in q.cores}, {i: K for i in range(6)},
                           {i: ("a", K, K) for i in range(5)}, "float32")
    assert plan.mps_ladder() is None
    plan = ContractionPlan(q.adjacency_table, 6, {c: (K,) * 4 for c in q.cores}, {i: K for i in range(6)},
                           {i: ("a", K, K) for i in range(6)}, "float64")
    assert plan.mps_ladder() is None


@pytest.mark.parametrize("n,K,B,nwarps", [(3, 3, 5, 2), (4, 3, 10, 3), (6, 3, 7, 1), (8, 3, 20, 4), (4, 2, 21, 2),
                                          (7, 2, 9, 5), (24, 3, 6, 2)])
def test_emulated_kernel_matches_oracle(emu, n, K, B, nwarps):
    graph = merged_graph(n, K)
    _, layer1, layer2 = ladder_of(graph, K)
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=n * 10 + K)
    scale = math.prod(m.scale for m in mxs)
    want = oc.forward(graph, cores, states, clone_mx(mxs)) / scale
    wl, wg = oc.loss_and_grads(graph, cores, states, clone_mx(mxs))
    # float64 yardstick: both float32 computations round differently over n qubits; ours must be
    # within 1e-5 or no worse than 3x the reference's own float32 error
    c64 = {k: v.double() for k, v in cores.items()}
    s64 = [s.double() for s in states]
    m64 = [oc.TNT(m.tensor.double(), m.scale, m.log_scale) for m in mxs]
    p64 = oc.forward(graph, c64, s64, clone_mx(m64)) / scale
    l64, g64 = oc.loss_and_grads(graph, c64, s64, clone_mx(m64))
    ms = [m.tensor.contiguous() for m in mxs]
    log_scale = sum(m.log_scale for m in mxs)
    v0, _, _ = run_emu(emu, K, n, cores, layer1, layer2, states, ms, B, 0, nwarps=nwarps)
    ref_err = ((want.double() - p64).abs() / p64.abs()).max().item()
    assert ((v0.double() - p64).abs() / p64.abs()).max().item() <= max(1e-5, 3 * ref_err)
    v1, loss, grads = run_emu(emu, K, n, cores, layer1, layer2, states, ms, B, 1, log_scale=log_scale, nwarps=nwarps)
    assert torch.equal(v0, v1)
    assert abs(float(loss) - float(l64)) <= max(1e-5 * abs(float(l64)), 3 * abs(float(wl) - float(l64)))
    for name, w, w64 in zip(names, wg, g64):
        ref_g = ((w.double() - w64).abs().max() / w64.abs().max()).item()
        assert ((grads[name].double() - w64).abs().max() / w64.abs().max()).item() <= max(1e-5, 3 * ref_g), name
    # autograd route: seed = d loss / d value of the same loss
    seed = (-1.0 / (B * v0.clamp_min(1e-10))).contiguous()
    _, _, g2 = run_emu(emu, K, n, cores, layer1, layer2, states, ms, B, 2, seed=seed, nwarps=nwarps)
    for name in names:
        assert (g2[name] - grads[name]).abs().max() <= 2e-6 * grads[name].abs().max()


def test_emulated_kernel_on_reference_fixture(emu):
    """Golden vectors produced by the real reference (oracle/make_golden.py)."""
    for name, K in (("merged6_k3_f32", 3), ("merged4_k2_f32", 2)):
        c = load_case(os.path.join(GOLDEN, name + ".npz"))
        n = len(c["mxs"])
        _, layer1, layer2 = ladder_of(c["graph"], K)
        mxs = fresh_mx(c)
        raw = [(m.tensor if isinstance(m, oc.TNT) else m).contiguous() for m in mxs]
        scale = math.prod(m.scale for m in mxs if isinstance(m, oc.TNT))
        log_scale = sum(m.log_scale for m in mxs if isinstance(m, oc.TNT))
        B = raw[0].shape[0]
        vals, loss, grads = run_emu(emu, K, n, c["cores"], layer1, layer2, c["states"], raw, B, 1, log_scale=log_scale)
        p64 = oc.forward(c["graph"], {k: v.double() for k, v in c["cores"].items()}, [s.double() for s in c["states"]],
                         [m.double() for m in raw])
        # the fixture's probabilities carry the reference's own float32 cancellation error: compare
        # both against float64 and demand we are not worse than 3x the reference
        ref_err = ((c["probabilities"].double() - p64 * scale).abs() / (p64 * scale).abs()).max().item()
        our_err = ((vals.double() - p64).abs() / p64.abs()).max().item()
        assert our_err <= max(3 * ref_err, 1e-5)
        if (p64 > 1e-8).all():
            assert abs(float(loss) - float(c["loss"])) <= 1e-4 * abs(float(c["loss"]))


def test_broadcast_measurement_and_tail_group(emu):
    """A (1,K,K) measurement shared by the batch (stride 0) and a batch that does not fill the last group."""
    n, K, B = 5, 3, 4
    graph = merged_graph(n, K)
    _, layer1, layer2 = ladder_of(graph, K)
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=5)
    ms = [m.tensor.contiguous() for m in mxs]
    eye = torch.eye(K).reshape(1, K, K).contiguous()
    ms[2] = eye
    strides = [K * K] * n
    strides[2] = 0
    full = [m if q != 2 else eye.expand(B, K, K).contiguous() for q, m in enumerate(ms)]
    want = oc.forward(graph, cores, states, full)
    v, _, _ = run_emu(emu, K, n, cores, layer1, layer2, states, ms, B, 0, strides=strides)
    assert ((v - want).abs() / want.abs()).max() < 1e-5
