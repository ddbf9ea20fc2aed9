"""Index bookkeeping parity: the symbolic greedy schedule must produce exactly the
einsum strings the reference hands to torch.einsum (north_star: "bit-exact
contraction ordering and index bookkeeping")."""
import json
import os

import pytest
import torch

import tneq_b200
from tneq_b200.contractor.greedy_plan import build_schedule
from tneq_b200.contractor.plan import signature_of
from oracle import qctn_oracle as oc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
EQ = json.load(open(os.path.join(GOLDEN, "equations.json")))
H = tneq_b200.QCTNHelper


@pytest.mark.parametrize("name", [k for k in EQ if k.startswith("strings_")])
def test_strings_against_reference_goldens(name):
    rec = EQ[name]
    q = tneq_b200.QCTN(rec["graph"])
    K, n = rec["K"], q.nqubits
    mi = {i: ("a" if rec["mode"] == "a" else "ab", K, K) for i in range(n)}
    sch = build_schedule(q.adjacency_table, n, {i: K for i in range(n)}, mi)
    assert sch.equations == rec["equations"]
    assert sch.batch == rec["mode"]


def _oracle_strings(graph, states, mxs):
    names, table, nq = oc.parse_graph(graph)
    torch.manual_seed(0)
    cores = oc.random_cores(table)
    log = []
    try:
        oc.greedy_contract(table, nq, cores, states, mxs, log=log)
    except RuntimeError:
        pass  # malformed final einsum of a disconnected network (reference defect D6)
    return [e for e, _ in log]


@pytest.mark.parametrize("kind,n,K", [("mps", 3, 2), ("mps", 7, 2), ("tree", 4, 2), ("tree", 5, 2), ("tree", 9, 2),
                                      ("wall", 4, 2), ("wall", 5, 2), ("mps", 4, 3)])
@pytest.mark.parametrize("variant", ["full", "ab", "none_mid", "none_ends", "no_state", "dict"])
def test_strings_against_oracle_for_input_variants(kind, n, K, variant):
    graph = H.generate_example_graph(n=n, graph_type=kind, dim_char=str(K))
    B = 2
    mxs = [torch.randn(B, K, K) for _ in range(n)]
    states = oc.unit_states(n, K)
    if variant == "ab":
        mxs = [torch.randn(B, 2, K, K) for _ in range(n)]
    elif variant == "none_mid":
        mxs[1] = None
    elif variant == "none_ends":
        mxs[0] = None
        mxs[-1] = None
    elif variant == "no_state":
        states = {q: s for q, s in enumerate(states) if q != 1}
    elif variant == "dict":
        mxs = {q: m for q, m in enumerate(mxs)}
    q = tneq_b200.QCTN(graph)
    sd, mi = signature_of(n, states, mxs)
    sch = build_schedule(q.adjacency_table, n, sd, mi)
    want = _oracle_strings(graph, states, mxs)
    assert sch.equations[: len(want)] == want and len(want) >= len(sch.equations) - 1


def test_merged_two_layer_network():
    q = tneq_b200.QCTN(H.generate_example_graph(n=4, graph_type="mps", dim_char="2"))
    m = tneq_b200.QCTN.merge(q, q)
    sch = build_schedule(m.adjacency_table, 4, {i: 2 for i in range(4)}, {i: ("a", 2, 2) for i in range(4)})
    assert sch.equations == EQ["merged4_k2_f32"]


def test_disconnected_network_keeps_reference_bookkeeping():
    """Two independent one-qubit cores: the per-qubit groups match the reference; the
    reference's trailing einsum is malformed (',->', defect D6) and is reproduced as
    bookkeeping, while raw_subs carry the true subscripts used by the device plan."""
    q = tneq_b200.QCTN("-2-a-2-\n-2-b-2-")
    sch = build_schedule(q.adjacency_table, 2, {0: 2, 1: 2}, {0: ("a", 2, 2), 1: ("a", 2, 2)})
    assert sch.equations == ["cd,c,ade,fe,f->a", "cd,c,ade,fe,f->a", ",->"]
    assert sch.steps[-1].raw_subs == ["a", "a"] and sch.steps[-1].raw_out == "a"
