"""The second-generation ladder kernel's arithmetic (csrc/tnq_ladder2_core.cuh) checked on the CPU.

tests/emu/ladder2_emu.cpp compiles the SAME phase functions and tile driver that the CUDA kernel uses
(csrc/tnq_ladder2.cu) with the 128 threads of a CTA executed one after the other, for every lane
geometry (R = 1, 2, 4, 8 row-block slots x 32/R samples per warp), so the row-block tables, the
shared-memory layouts, the fused phases, the reverse sweep with its recomputation, the flush
machinery and the per-tile gradient slices are validated against the oracle and the reference
fixtures without a GPU.  Test infrastructure: nothing in the product can reach the emulation.
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
from oracle import qctn_oracle as oc
from helpers import well_conditioned_case, clone_mx
from test_oracle_golden import load_case, fresh_mx, GOLDEN
from test_ladder_emu import merged_graph, ladder_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "quantum_circuits_symmetry_breaking_based_on_tneq-qc_b200", "csrc")


@pytest.fixture(scope="module")
def emu2(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu2") / "ladder2_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                    os.path.join(ROOT, "tests", "emu", "ladder2_emu.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def run_emu2(lib, R, n, cores, layer1, layer2, states, ms, B, mode, seed=None, log_scale=0.0, strides=None, warps=4):
    K = 3
    arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    ca = [cores[k].contiguous() for k in layer1]
    cx = [cores[k].contiguous() for k in layer2]
    ga, gx = [torch.zeros_like(c) for c in ca], [torch.zeros_like(c) for c in cx]
    vals, loss = torch.zeros(B), torch.zeros(1)
    sd = seed if seed is not None else torch.zeros(B)
    st = (ctypes.c_longlong * n)(*(strides or [K * K] * n))
    rc = lib.ladder2_emu(R, warps, n, arr(ca), arr(cx), arr(states), arr(ms), st, ctypes.c_longlong(B), mode,
                         ctypes.c_void_p(sd.data_ptr()), ctypes.c_void_p(vals.data_ptr()),
                         ctypes.c_void_p(loss.data_ptr()), arr(ga), arr(gx), ctypes.c_double(log_scale))
    assert rc == 0
    grads = dict(zip(layer1, ga))
    grads.update(zip(layer2, gx))
    return vals, loss[0], grads


@pytest.mark.parametrize("R", [1, 2, 4, 8])
def test_row_block_tables(emu2, R):
    """Every row block exactly once, rb_of / uslot_of inverse of each other, and the bank rule: the row
    blocks a warp works on at the same time sit at positions that are equal or distinct modulo R."""
    assert emu2.ladder2_check_tables(R) == 0


@pytest.mark.parametrize("R,n,B", [(1, 3, 5), (2, 3, 5), (2, 4, 21), (1, 6, 12), (2, 6, 11), (4, 6, 9), (8, 6, 5),
                                   (2, 24, 5), (4, 8, 8)])
def test_emulated_kernel_matches_oracle(emu2, R, n, B):
    K = 3
    graph = merged_graph(n, K)
    _, layer1, layer2 = ladder_of(graph, K)
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=n * 10 + R)
    scale = math.prod(m.scale for m in mxs)
    log_scale = sum(m.log_scale for m in mxs)
    want = oc.forward(graph, cores, states, clone_mx(mxs))
    raw = [m.tensor.contiguous() for m in mxs]
    vals, _, _ = run_emu2(emu2, R, n, cores, layer1, layer2, states, raw, B, 0)
    assert (vals * scale - want).abs().max() <= 2e-5 * want.abs().max()
    c64 = {k: v.double() for k, v in cores.items()}
    tl, tg = oc.loss_and_grads(graph, c64, [s.double() for s in states],
                               [oc.TNT(m.tensor.double(), m.scale, m.log_scale) for m in clone_mx(mxs)])
    wl, wg = oc.loss_and_grads(graph, cores, states, clone_mx(mxs))
    vals1, loss, grads = run_emu2(emu2, R, n, cores, layer1, layer2, states, raw, B, 1, log_scale=log_scale)
    assert torch.equal(vals1, vals)
    assert abs(loss.item() - tl.item()) <= max(1e-5 * abs(tl.item()), 8 * abs(wl.item() - tl.item()))
    for name, t, w in zip(names, tg, wg):
        err = ((grads[name].double() - t).abs().max() / t.abs().max()).item()
        ref = ((w.double() - t).abs().max() / t.abs().max()).item()
        assert err <= max(1e-5, 8 * ref), (name, err, ref)
    # MODE 2: the reverse sweep seeded with d loss / d value reproduces the fused gradients
    v = vals.clone().requires_grad_(True)
    (-(torch.log(torch.clamp(v, min=1e-10)) + log_scale).mean()).backward()
    _, _, grads2 = run_emu2(emu2, R, n, cores, layer1, layer2, states, raw, B, 2, seed=v.grad.contiguous())
    for name in names:
        assert (grads2[name] - grads[name]).abs().max() <= 1e-6 * grads[name].abs().max() + 1e-12


def test_emulated_kernel_matches_reference_fixture(emu2):
    """tests/golden/merged6_k3_f32: numbers produced by the reference itself."""
    c = load_case(os.path.join(GOLDEN, "merged6_k3_f32.npz"))
    K, graph = c["K"], c["graph"]
    _, layer1, layer2 = ladder_of(graph, K)
    mxs = fresh_mx(c)
    B = c["probabilities"].shape[0]
    n = len(c["states"])
    scale = math.prod(m.scale for m in mxs)
    log_scale = sum(m.log_scale for m in mxs)
    raw = [m.tensor.contiguous() for m in mxs]
    for R in (1, 2, 4, 8):
        vals, loss, grads = run_emu2(emu2, R, n, c["cores"], layer1, layer2, c["states"], raw, B, 1, log_scale=log_scale)
        assert (vals * scale - c["probabilities"]).abs().max() <= 2e-5 * c["probabilities"].abs().max()
        assert abs(loss.item() - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
        for name, w in zip(c["names"], c["grads"]):
            assert (grads[name] - w).abs().max() <= 5e-3 * w.abs().max()


@pytest.mark.parametrize("R,n,B", [(2, 5, 21), (4, 6, 11), (2, 3, 7)])
def test_eight_warp_variant_matches_four_warp_variant(emu2, R, n, B):
    """The small-batch geometry (8 warps per CTA: half the row blocks per warp) computes what the 4-warp CTA
    computes: values bit-identical, gradients equal up to the order of the partial sums."""
    K = 3
    graph = merged_graph(n, K)
    _, layer1, layer2 = ladder_of(graph, K)
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=n)
    raw = [m.tensor.contiguous() for m in mxs]
    v4, l4, g4 = run_emu2(emu2, R, n, cores, layer1, layer2, states, raw, B, 1, warps=4)
    v8, l8, g8 = run_emu2(emu2, R, n, cores, layer1, layer2, states, raw, B, 1, warps=8)
    assert torch.equal(v4, v8)
    assert abs(l4.item() - l8.item()) <= 1e-6 * abs(l4.item())
    for k in names:
        assert (g4[k] - g8[k]).abs().max() <= 2e-5 * g4[k].abs().max() + 1e-12


def test_geometries_agree_and_ragged_batches(emu2):
    """All lane geometries compute the same per-sample arithmetic (values bit-identical); a batch that
    is not a multiple of the tile, broadcast (stride 0) measurements."""
    K, n, B = 3, 5, 37
    graph = merged_graph(n, K)
    _, layer1, layer2 = ladder_of(graph, K)
    names, table, nq, cores, states, mxs = well_conditioned_case(graph, K, B, "float32", seed=5)
    raw = [m.tensor.contiguous() for m in mxs]
    base = None
    for R in (1, 2, 4, 8):
        vals, loss, grads = run_emu2(emu2, R, n, cores, layer1, layer2, states, raw, B, 1)
        if base is None:
            base = (vals, loss, grads)
            continue
        assert torch.equal(vals, base[0])
        assert abs(loss.item() - base[1].item()) <= 1e-6 * abs(base[1].item())
        for k in names:
            assert (grads[k] - base[2][k]).abs().max() <= 2e-5 * base[2][k].abs().max()
    # identity measurements broadcast with stride 0: orthogonal cores => value 1
    eye = [torch.eye(K).reshape(1, K, K).contiguous() for _ in range(n)]
    vals, _, _ = run_emu2(emu2, 2, n, cores, layer1, layer2, states, eye, 11, 0, strides=[0] * n)
    assert (vals - 1).abs().max() < 1e-5
