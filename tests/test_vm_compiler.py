"""Plan compiler end to end on CPU: schedule -> pairwise graph -> reverse mode -> device program,
executed by the numpy emulator of the program format and compared with the reference's own
outputs (tests/golden, produced by /root/reference) and with the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import tneq_b200
from tneq_b200.contractor.plan import ContractionPlan, signature_of
from oracle import qctn_oracle as oc
import vm_emulator as em
from test_oracle_golden import FILES, load_case, fresh_mx


def to_real(t):
    t = t.detach()
    return (torch.view_as_real(t) if t.is_complex() else t).contiguous().numpy()


def marshal(prog, cores, states, mxs, nsamples, nb, gradseed=None):
    ins = []
    for s in prog.inputs:
        kind, key = s.key
        if kind == "core":
            x = to_real(cores[key]).reshape(-1)
        elif kind == "state":
            x = to_real(states[key]).reshape(-1)
        elif kind == "mx":
            m = oc._raw(mxs[key])
            x = to_real(m)
            if m.ndim == 3 and nb == 2:
                x = np.repeat(x.reshape(m.shape[0], 1, -1), 2, axis=1)
            x = x.reshape(nsamples, -1)
        elif kind == "gradseed":
            x = gradseed
        ins.append(x)
    return ins


def plan_for(graph, states, mxs, dtype):
    q = tneq_b200.QCTN(graph)
    sd, mi = signature_of(q.nqubits, states, mxs)
    return ContractionPlan(q.adjacency_table, q.nqubits, {c: q.core_shape(c) for c in q.cores}, sd, mi, dtype)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_device_program_matches_reference_fixture(path):
    """float64 emulation of the program on the fixture's float32 inputs vs the reference's float32
    outputs: agreement to float32 round-off, for values, loss and every core gradient."""
    c = load_case(path)
    plan = plan_for(c["graph"], c["states"], c["mxs"], c["dtype"])
    assert plan.equations == [e for e in plan.schedule.equations]
    B = c["probabilities"].shape[0]
    pf = plan.program("fwd")
    raw = em.run(pf.to_blob(), marshal(pf, c["cores"], c["states"], c["mxs"], B, 1), B, dtype=np.float64)[0]
    if plan.complex_mode:
        val = raw[:, 0] ** 2 + raw[:, 1] ** 2
    else:
        val = raw[:, 0]
    scale = 1.0
    for m in c["mxs"]:
        if isinstance(m, oc.TNT):
            scale *= m.scale
    if plan.complex_mode:
        scale = scale ** 2
    want = c["probabilities"].double().numpy()
    assert np.abs(val * scale - want).max() <= 2e-5 * np.abs(want).max()

    pt = plan.program("train")
    lscale = sum(m.log_scale for m in c["mxs"] if isinstance(m, oc.TNT))
    outs = em.run(pt.to_blob(), marshal(pt, c["cores"], c["states"], c["mxs"], B, 1), B,
                  scalars=(lscale, 1.0 / B), dtype=np.float64)
    loss = outs[pt.output_index(("loss", 0))][0, 0]
    # the fixture's float32 loss/grads carry the reference's own round-off (ill-conditioned
    # samples included), hence the looser bound on gradients
    assert abs(loss - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    for name, gw in zip(c["names"], c["grads"]):
        gg = outs[pt.output_index(("grad", "core", name))].reshape(-1)
        gw = to_real(gw).reshape(-1).astype(np.float64)
        assert np.abs(gg - gw).max() <= 5e-3 * np.abs(gw).max() + 1e-12


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
@pytest.mark.parametrize("kind,n,K,mode", [("mps", 5, 3, "a"), ("tree", 6, 2, "a"), ("wall", 4, 2, "a"),
                                           ("merged", 4, 2, "a"), ("mps", 5, 2, "ab")])
def test_device_program_matches_oracle_in_double(kind, n, K, mode, dtype):
    H = tneq_b200.QCTNHelper
    if kind == "merged":
        q = tneq_b200.QCTN(H.generate_example_graph(n=n, graph_type="mps", dim_char=str(K)))
        graph = tneq_b200.QCTN.merge(q, q).graph
    else:
        graph = H.generate_example_graph(n=n, graph_type=kind, dim_char=str(K))
    from helpers import make_case, clone_mx
    names, table, nq, cores, states, mxs = make_case(graph, K, 7, dtype, tnt=True, mode=mode)
    plan = plan_for(graph, states, mxs, dtype)
    nb, NS = plan.nb, 7 * plan.nb
    want = to_real(oc._raw(oc.greedy_contract(table, nq, cores, states, clone_mx(mxs)))).reshape(NS, -1)
    pf = plan.program("fwd")
    got = em.run(pf.to_blob(), marshal(pf, cores, states, mxs, NS, nb), NS, dtype=np.float64)[0]
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    if mode == "ab":
        return
    wl, wg = oc.loss_and_grads(graph, cores, states, clone_mx(mxs))
    lscale = sum(m.log_scale for m in mxs if isinstance(m, oc.TNT))
    for prog_mode in ("train", "bwd"):
        p = plan.program(prog_mode)
        if prog_mode == "train":
            outs = em.run(p.to_blob(), marshal(p, cores, states, mxs, NS, nb), NS, scalars=(lscale, 1.0 / NS),
                          dtype=np.float64)
            assert abs(outs[p.output_index(("loss", 0))][0, 0] - float(wl)) <= 1e-12 * abs(float(wl))
        else:
            # seed = d loss / d result, as torch.autograd would hand it to the compute function
            res = oc._raw(oc.greedy_contract(table, nq, cores, states, clone_mx(mxs)))
            resr = res.detach().clone().requires_grad_(True)
            val = oc.abs_square(resr)
            (-(torch.log(torch.clamp(val, min=1e-10)) + lscale).mean()).backward()
            seed = to_real(resr.grad).reshape(NS, -1)
            outs = em.run(p.to_blob(), marshal(p, cores, states, mxs, NS, nb, gradseed=seed), NS, dtype=np.float64)
        for name, gw in zip(names, wg):
            gg = outs[p.output_index(("grad", "core", name))].reshape(-1)
            gw = to_real(gw).reshape(-1)
            assert np.abs(gg - gw).max() <= 1e-10 * np.abs(gw).max() + 1e-300


def test_disconnected_network_is_the_product_of_its_parts():
    """The reference crashes here (defect D6); the device plan returns the true value."""
    graph = "-2-a-2-\n-2-b-2-"
    from helpers import make_case
    names, table, nq, cores, states, mxs = make_case(graph, 2, 5, "float64", tnt=False)
    plan = plan_for(graph, states, mxs, "float64")
    pf = plan.program("fwd")
    got = em.run(pf.to_blob(), marshal(pf, cores, states, mxs, 5, 1), 5, dtype=np.float64)[0][:, 0]
    want = np.ones(5)
    for q, c in enumerate(names):
        u = cores[c].numpy()
        psi = u[1, :] if False else states[q].numpy() @ u          # <s|U
        want = want * np.einsum("i,bij,j->b", psi, mxs[q].numpy(), psi)
    assert np.allclose(got, want, rtol=1e-12)


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_right_qctn_given_as_second_network(dtype):
    """`right_qctn=<QCTN>` (engine_siamese.py:304,390; the 'qctn' branch of the greedy sweep): the
    right-hand copy is a second set of cores, used as given (no conjugation, no transposition).
    Values and the gradients of both core sets through the device program vs the oracle; where the
    reference checkout exists, the oracle's numbers are first checked against the reference's own."""
    from helpers import make_case, clone_mx
    K, n, B = 2, 4, 6
    graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
    names, table, nq, cores, states, mxs = make_case(graph, K, B, dtype, tnt=False)
    torch.manual_seed(5)
    rcores = {k: v + 0.2 * torch.randn_like(v) for k, v in cores.items()}
    want = oc.greedy_contract(table, nq, cores, states, clone_mx(mxs), right="qctn", right_table=table, right_cores=rcores)
    from oracle import ref_harness as rh
    if rh.available():
        import contextlib, io
        ns = rh.load()
        with contextlib.redirect_stdout(io.StringIO()):
            be = ns.BackendFactory.create_backend("pytorch", device="cpu", dtype=dtype)
            eng = ns.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
            ql, qr = ns.QCTN(graph, backend=be), ns.QCTN(graph, backend=be)
            for k in names:
                ql.cores_weights[k], qr.cores_weights[k] = cores[k], rcores[k]
            ref = eng.contract_with_compiled_strategy(ql, states, clone_mx(mxs), right_qctn=qr)
        assert torch.equal(ref, oc.abs_square(want))
    q = tneq_b200.QCTN(graph)
    sd, mi = signature_of(q.nqubits, states, mxs)
    shapes = {c: q.core_shape(c) for c in q.cores}
    plan = ContractionPlan(q.adjacency_table, q.nqubits, shapes, sd, mi, dtype, right="qctn",
                           right_table=q.adjacency_table, right_core_shapes=shapes)

    slot_elems = {}

    def marshal2(prog, gradseed=None):
        ins = []
        for s in prog.inputs:
            slot_elems[s.key] = s.elems
            kind, key = s.key
            if kind == "rcore":
                ins.append(to_real(rcores[key]).reshape(-1))
            else:
                ins.append(marshal_one(kind, key, gradseed))
        return ins

    def marshal_one(kind, key, gradseed):
        if kind == "core":
            return to_real(cores[key]).reshape(-1)
        if kind == "state":
            return to_real(states[key]).reshape(-1)
        if kind == "mx":
            return to_real(oc._raw(mxs[key])).reshape(B, -1)
        if kind == "ones":
            return to_real(torch.ones(slot_elems[(kind, key)] // (2 if plan.complex_mode else 1), dtype=cores[names[0]].dtype)).reshape(-1)
        return gradseed

    pf = plan.program("fwd")
    got = em.run(pf.to_blob(), marshal2(pf), B, dtype=np.float64)[0]
    assert np.abs(got - to_real(want).reshape(B, -1)).max() <= 1e-12 * np.abs(to_real(want)).max()
    # gradients of both core sets: seed = d sum(value) / d result
    cl = {k: v.clone().requires_grad_(True) for k, v in cores.items()}
    cr = {k: v.clone().requires_grad_(True) for k, v in rcores.items()}
    res = oc.greedy_contract(table, nq, cl, states, clone_mx(mxs), right="qctn", right_table=table, right_cores=cr)
    resr = res.detach().clone().requires_grad_(True)
    oc.abs_square(resr).sum().backward()
    oc.abs_square(res).sum().backward()
    pb = plan.program("bwd")
    outs = em.run(pb.to_blob(), marshal2(pb, gradseed=to_real(resr.grad).reshape(B, -1)), B, dtype=np.float64)
    for kind, src in (("core", cl), ("rcore", cr)):
        for name in names:
            gg = outs[pb.output_index(("grad", kind, name))].reshape(-1)
            gw = to_real(src[name].grad).reshape(-1)
            assert np.abs(gg - gw).max() <= 1e-10 * np.abs(gw).max() + 1e-300, (kind, name)
