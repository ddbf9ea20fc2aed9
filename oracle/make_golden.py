"""Generate tests/golden/* from the REAL reference (run in the build container only).

    python oracle/make_golden.py

For every case the unmodified reference (/root/reference, imported through
oracle/ref_harness.py) computes probabilities, loss, per-core gradients and the
per-qubit einsum strings on CPU; this script
  1. asserts that oracle/qctn_oracle.py reproduces all of it BIT-IDENTICALLY
     (same einsum strings, same values), which is what pins the oracle, and
  2. writes the inputs and the reference's outputs as small .npz fixtures plus
     equations.json, so the pin can be re-checked where /root/reference does
     not exist (the GPU box) and so the CUDA path can be compared with numbers
     that came from the reference itself.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import qctn_oracle as oc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, graph kind, n, K, B, dtype, TNTensor measurements
    ("mps6_k3_f32", "mps", 6, 3, 8, "float32", True),
    ("mps6_k3_f32_plain", "mps", 6, 3, 8, "float32", False),
    ("mps16_k3_f32", "mps", 16, 3, 6, "float32", True),
    ("mps5_k2_f64", "mps", 5, 2, 7, "float64", True),
    ("mps6_k3_c64", "mps", 6, 3, 5, "complex64", True),
    ("tree6_k2_f32", "tree", 6, 2, 9, "float32", True),
    ("tree7_k3_f32", "tree", 7, 3, 4, "float32", True),
    ("merged4_k2_f32", "merged", 4, 2, 6, "float32", True),
    ("merged6_k3_f32", "merged", 6, 3, 3, "float32", True),
    ("merged4_k2_c64", "merged", 4, 2, 4, "complex64", True),
]


def graph_of(ns, kind, n, K):
    H = ns.QCTNHelper
    if kind == "merged":
        be, _ = rh.make_engine()
        with rh.quiet():
            q = ns.QCTN(H.generate_example_graph(n=n, graph_type="mps", dim_char=str(K)), backend=be)
            return ns.QCTN.merge(q, q).graph
    return H.generate_example_graph(n=n, graph_type=kind, dim_char=str(K))


def to_np(t):
    t = t.detach()
    return t.numpy()


def main():
    ns = rh.load()
    os.makedirs(GOLDEN, exist_ok=True)
    equations = {}
    for name, kind, n, K, B, dtype, tnt in CASES:
        torch.manual_seed(1234)
        td = getattr(torch, dtype)
        graph = graph_of(ns, kind, n, K)
        names, table, nq = oc.parse_graph(graph)
        cores = oc.random_cores(table, td)
        torch.manual_seed(42)
        x = torch.randn(B, nq)
        be, eng = rh.make_engine(dtype, K)
        ret = "TNTensor" if tnt else "tensor"
        states = oc.unit_states(nq, K, td)

        mref, _ = eng.generate_data(x, K=K, ret_type=ret)
        log_ref, log_or = [], []
        p_ref = rh.ref_forward(graph, cores, states, mref, dtype, log_ref)
        mor, _ = oc.generate_data(x, K, td, ret)
        p_or = oc.forward(graph, cores, states, mor, log=log_or)
        assert log_ref == log_or, f"{name}: einsum bookkeeping differs"
        assert torch.equal(p_ref, p_or), f"{name}: oracle forward is not bit-identical to the reference"

        mref, _ = eng.generate_data(x, K=K, ret_type=ret)
        l_ref, g_ref = rh.ref_loss_and_grads(graph, cores, states, mref, dtype)
        mor, _ = oc.generate_data(x, K, td, ret)
        l_or, g_or = oc.loss_and_grads(graph, cores, states, mor)
        assert torch.equal(l_ref, l_or), f"{name}: oracle loss is not bit-identical"
        assert all(torch.equal(a, b) for a, b in zip(g_ref, g_or)), f"{name}: oracle grads are not bit-identical"

        mor, _ = oc.generate_data(x, K, td, ret)
        blob = {"graph": np.array(graph), "dtype": np.array(dtype), "K": np.array(K), "x": to_np(x),
                "probabilities": to_np(p_ref), "loss": to_np(l_ref)}
        for c in names:
            blob[f"core_{c}"] = to_np(cores[c])
        for c, gr in zip(names, g_ref):
            blob[f"grad_{c}"] = to_np(gr)
        for q, m in enumerate(mor):
            blob[f"mx_{q}"] = to_np(oc._raw(m))
            blob[f"mx_scale_{q}"] = np.array([m.scale, m.log_scale] if isinstance(m, oc.TNT) else [1.0, 0.0])
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)
        equations[name] = [e for e, _ in log_ref]
        print(f"{name}: ok  loss={float(l_ref):.6f}  steps={len(log_ref)}")

    # bookkeeping-only goldens (strings), including the headline 24-qubit two-layer network
    for name, kind, n, K, mode in [("mps24_merged_k3", "merged", 24, 3, "a"), ("mps16_k3", "mps", 16, 3, "a"),
                                   ("tree8_k2", "tree", 8, 2, "a"), ("mps6_k3_ab", "mps", 6, 3, "ab"),
                                   ("wall4_k2", "wall", 4, 2, "a")]:
        graph = graph_of(ns, kind, n, K)
        names, table, nq = oc.parse_graph(graph)
        torch.manual_seed(0)
        cores = oc.random_cores(table)
        B = 2
        mx = [torch.randn(B, K, K) if mode == "a" else torch.randn(B, 2, K, K) for _ in range(nq)]
        log = []
        rh.ref_forward(graph, cores, oc.unit_states(nq, K), mx, "float32", log)
        equations["strings_" + name] = {"graph": graph, "mode": mode, "K": K, "equations": [e for e, _ in log]}
        print(f"strings_{name}: {len(log)} steps")
    with open(os.path.join(GOLDEN, "equations.json"), "w") as f:
        json.dump(equations, f, indent=1)
    sgdg_golden(ns)


def sgdg_golden(ns):
    """SGDG / Cayley optimizer step (backend_pytorch.py:200-268 optimizer_update, :349-468 _sgdg_step):
    three consecutive steps of the REAL reference on plain tensors, with and without momentum, in
    float32 and float64; asserts oc.sgdg_step bit-identical and writes tests/golden/sgdg_*.npz.
    The 1 % QR retraction draws from Python's `random`: both sides are seeded identically, and the
    seed is chosen so that one of the draws fires inside the three steps."""
    import random
    for name, dtype, momentum, seed in [("sgdg_f32", "float32", 0.0, 7), ("sgdg_f32_mom", "float32", 0.9, 7),
                                        ("sgdg_f64", "float64", 0.5, 3)]:
        td = getattr(torch, dtype)
        be, _ = rh.make_engine(dtype, 3)
        torch.manual_seed(99)
        shapes = [(3, 3, 3, 3), (3, 3, 3, 3), (2, 2, 2, 2), (3, 9), (4, 4, 4, 4)]
        params0 = [torch.linalg.qr(torch.randn(s[1], s[0], dtype=td))[0].T.contiguous() if len(s) == 2
                   else oc.init_random_core(s, td) for s in shapes]       # (3, 9): a rectangular edge core
        grads = [[0.3 * torch.randn(s, dtype=td) for s in shapes] for _ in range(3)]
        hp = {"learning_rate": 0.05, "momentum": momentum, "stiefel": True}
        # find a Python-random seed that fires the retraction at least once in 3 x 5 draws
        fire = None
        for cand in range(seed, seed + 2000):
            r = random.Random(cand)
            if any(r.randint(1, 101) == 1 for _ in range(15)):
                fire = cand
                break
        random.seed(fire)
        p_ref, st_ref = [p.clone() for p in params0], {}
        ref_steps = []
        for g in grads:
            p_ref, st_ref = be.optimizer_update(list(p_ref), [x.clone() for x in g], st_ref, "sgdg", hp)
            p_ref = [p.detach() for p in p_ref]
            ref_steps.append([p.clone() for p in p_ref])
        rng = random.Random(fire)
        p_or, st_or = [p.clone() for p in params0], {}
        for k, g in enumerate(grads):
            p_or, st_or = oc.sgdg_step(p_or, g, st_or, lr=0.05, momentum=momentum, stiefel=True, rng=rng)
            assert all(torch.equal(a, b) for a, b in zip(p_or, ref_steps[k])), f"{name}: SGDG step {k} not bit-identical"
        blob = {"dtype": np.array(dtype), "momentum": np.array(momentum), "lr": np.array(0.05), "rng_seed": np.array(fire)}
        for i, p in enumerate(params0):
            blob[f"param_{i}"] = to_np(p)
        for k in range(3):
            for i in range(len(shapes)):
                blob[f"grad_{k}_{i}"] = to_np(grads[k][i])
                blob[f"step_{k}_{i}"] = to_np(ref_steps[k][i])
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **blob)
        print(f"{name}: ok (python-random seed {fire}, 3 steps bit-identical)")


if __name__ == "__main__":
    main()
