"""The oracle against fixtures produced by the REAL reference (oracle/make_golden.py).

In the build container make_golden.py asserts bit-identity; here (any machine)
the same numbers must be reproduced to float round-off, which keeps the oracle
pinned where /root/reference does not exist.
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import qctn_oracle as oc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FILES = sorted(f for f in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(f).startswith("sgdg_"))
SGDG_FILES = sorted(glob.glob(os.path.join(GOLDEN, "sgdg_*.npz")))


def load_case(path):
    z = np.load(path)
    graph, dtype, K = str(z["graph"]), str(z["dtype"]), int(z["K"])
    names, table, nq = oc.parse_graph(graph)
    cores = {c: torch.from_numpy(z[f"core_{c}"]) for c in names}
    mxs = []
    for q in range(nq):
        m = torch.from_numpy(z[f"mx_{q}"].copy())
        sc, ls = z[f"mx_scale_{q}"]
        mxs.append(oc.TNT(m, float(sc), float(ls)) if (sc, ls) != (1.0, 0.0) else m)
    states = oc.unit_states(nq, K, getattr(torch, dtype))
    grads = [torch.from_numpy(z[f"grad_{c}"]) for c in names]
    return dict(graph=graph, dtype=dtype, K=K, names=names, cores=cores, mxs=mxs, states=states,
                probabilities=torch.from_numpy(z["probabilities"]), loss=torch.from_numpy(z["loss"]), grads=grads)


def fresh_mx(case):
    return [oc.TNT(m.tensor.clone(), m.scale, m.log_scale) if isinstance(m, oc.TNT) else m.clone() for m in case["mxs"]]


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_reference_outputs(path):
    c = load_case(path)
    rt = 1e-6 if c["dtype"] in ("float32", "complex64") else 1e-13
    got = oc.forward(c["graph"], c["cores"], c["states"], fresh_mx(c))
    assert torch.allclose(got, c["probabilities"], rtol=rt * 10, atol=0)
    loss, grads = oc.loss_and_grads(c["graph"], c["cores"], c["states"], fresh_mx(c))
    assert abs(float(loss) - float(c["loss"])) <= rt * abs(float(c["loss"]))
    for g, w in zip(grads, c["grads"]):
        assert (g - w).abs().max() <= rt * 10 * w.abs().max()


def test_oracle_einsum_strings_match_reference():
    eq = json.load(open(os.path.join(GOLDEN, "equations.json")))
    for path in FILES:
        name = os.path.basename(path)[:-4]
        c = load_case(path)
        log = []
        oc.forward(c["graph"], c["cores"], c["states"], fresh_mx(c), log=log)
        assert [e for e, _ in log] == eq[name]


def test_survey_golden_strings():
    """The strings quoted in SURVEY.md 3.1 (captured from the reference)."""
    eq = json.load(open(os.path.join(GOLDEN, "equations.json")))
    assert eq["mps6_k3_f32"] == ["cdef,c,aeg,higj,h,d,i->ajf"] + ["cdef,aeg,higj,ahc,d,i->ajf"] * 4 + ["acd,adc->a"]
    assert eq["merged4_k2_f32"] == ["cdef,eghi,c,ahj,klmn,mojp,k,d,l->agnpfio",
                                    "cdef,ghij,aik,lmno,pqkr,aelpcgn,d,m->ahorfjq",
                                    "cdef,gfhi,ahj,klmn,onjp,aekocgm,d,l->api", "acd,adc->a"]
    assert eq["tree6_k2_f32"][:3] == ["cdef,c,aeg,higj,h->adjfi", "cdef,c,agh,ijkl,i,aehgk->adlfj",
                                      "cdef,c,agh,ijkl,i,aehgk,d,j->alf"]


def test_known_answers():
    """KAT-1 (orthogonal cores + identity measurements -> 1) and KAT-4 (autograd vs
    central finite differences in float64), SURVEY 8(c)."""
    K, n, B = 3, 5, 4
    graph = oc_graph = "".join(f"-{K}-" + ("" if False else "") for _ in range(0))  # placeholder, replaced below
    import tneq_b200
    graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
    names, table, nq = oc.parse_graph(graph)
    torch.manual_seed(3)
    cores = oc.random_cores(table, torch.float64)
    states = oc.unit_states(nq, K, torch.float64)
    eye = [torch.eye(K, dtype=torch.float64).expand(B, K, K) for _ in range(nq)]
    assert torch.allclose(oc.forward(graph, cores, states, eye), torch.ones(B, dtype=torch.float64), atol=1e-12)
    x = torch.randn(B, nq)
    mx, _ = oc.generate_data(x, K, torch.float64)
    mx = [m * 1e3 for m in mx]  # keep values above the 1e-10 clamp
    loss, grads = oc.loss_and_grads(graph, cores, states, mx)
    c0 = names[1]
    idx = (1, 0, 2, 1)
    h = 1e-6
    vals = []
    for sgn in (+1, -1):
        pert = {k: v.clone() for k, v in cores.items()}
        pert[c0][idx] += sgn * h
        vals.append(float(oc.loss_and_grads(graph, pert, states, mx)[0]))
    fd = (vals[0] - vals[1]) / (2 * h)
    assert abs(fd - float(grads[1][idx])) < 1e-6 * max(1.0, abs(fd))


def load_sgdg(path):
    z = np.load(path)
    n = len([k for k in z.files if k.startswith("param_")])
    params = [torch.from_numpy(z[f"param_{i}"]) for i in range(n)]
    grads = [[torch.from_numpy(z[f"grad_{k}_{i}"]) for i in range(n)] for k in range(3)]
    steps = [[torch.from_numpy(z[f"step_{k}_{i}"]) for i in range(n)] for k in range(3)]
    return params, grads, steps, float(z["lr"]), float(z["momentum"]), int(z["rng_seed"]), str(z["dtype"])


@pytest.mark.parametrize("path", SGDG_FILES, ids=[os.path.basename(f)[:-4] for f in SGDG_FILES])
def test_oracle_sgdg_reproduces_reference_steps(path):
    """oc.sgdg_step against three consecutive steps of the REAL reference's optimizer_update('sgdg')
    (backend_pytorch.py:200-268, 349-468; fixtures written by oracle/make_golden.py, which asserts
    bit-identity in the build container).  The Python-random seed fires the 1 % QR retraction."""
    import random
    params, grads, steps, lr, momentum, seed, dtype = load_sgdg(path)
    rt = 1e-5 if dtype == "float32" else 1e-12
    rng = random.Random(seed)
    state = {}
    for k in range(3):
        params, state = oc.sgdg_step(params, grads[k], state, lr=lr, momentum=momentum, stiefel=True, rng=rng)
        for a, b in zip(params, steps[k]):
            assert a.shape == b.shape and (a - b).abs().max() <= rt * b.abs().max()
