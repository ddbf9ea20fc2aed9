"""The algorithm of tnq_mps_chain_sample (csrc/tnq_chain.cu), restated in numpy float64 exactly as the kernel computes it
(left environment advanced by one chain step per qubit, shared right environments built from the back of the chain and
folded with the core, value linear in the measurement matrix of the qubit being sampled, running-sum CDF, #(cdf < u)
clamped to G - 2, the reference's interpolation), against the REFERENCE's procedure -- EngineSiamese.sample,
tneq_qc/core/engine_siamese.py:740-915: one full contraction per qubit and grid point -- evaluated with the oracle on the
CPU from the same uniform numbers.  No GPU: this pins the mathematics; tests/test_gpu_parity.py pins the kernel."""
import numpy as np
import pytest
import torch

import tneq_b200
from oracle import qctn_oracle as oc


def _phi(y, K):
    w = oc.hermite_weights(K)[:K]
    g = np.sqrt(np.exp(-(y * y) / 2.0))
    H = [1.0, y]
    for i in range(2, K):
        H.append(y * H[i - 1] - (i - 1) * H[i - 2])
    return np.array([w[i] * g * H[i] for i in range(K)])


def _interp(cdf, u, grid):
    G = len(grid)
    idx = min(int((cdf < u).sum()), G - 2)
    c0, c1 = cdf[idx], cdf[idx + 1]
    return grid[idx] + (u - c0) / (c1 - c0 + 1e-10) * (grid[idx + 1] - grid[idx])


def prefix_sample(cores, states, u, grid, K):
    """numpy restatement of tnq_chain_sample_kernel; cores[q][c,d,e,f], states[q][K], u[S,n]"""
    n = len(states)
    S = u.shape[0]
    Ls = [np.einsum("cdef,d->cef", cores[q], states[q + 1]) for q in range(n - 1)]
    mg = np.stack([np.outer(_phi(x, K), _phi(x, K)) for x in grid])            # [G][e][g]
    R = np.eye(K)                                                               # R_{n-1}[j][f]
    LR = [None] * (n - 1)
    for q in range(n - 2, -1, -1):
        LR[q] = np.einsum("hgj,jf->hgf", Ls[q], R)
        R = np.einsum("cef,hef->hc", Ls[q], LR[q])
    out = np.zeros((S, n))
    for s in range(S):
        env = np.outer(states[0], states[0])
        for q in range(n):
            if q < n - 1:
                T1 = np.einsum("hc,cef->hef", env, Ls[q])
                C = np.einsum("hef,hgf->eg", T1, LR[q])
            else:
                C = env.T
            dens = np.maximum(np.einsum("eg,ieg->i", C, mg), 0.0)
            cdf = np.cumsum(dens)
            cdf = cdf / (cdf[-1] + 1e-10)
            y = _interp(cdf, u[s, q], grid)
            out[s, q] = y
            if q < n - 1:
                ph = _phi(y, K)
                M = np.outer(ph, ph)
                env = np.einsum("cef,eg,hgj,hc->jf", Ls[q], M, Ls[q], env)     # the chain step of tnq_chain.cu
    return out


def reference_sample(graph, cores, states, u, grid, K):
    """engine_siamese.py:802-905 with the oracle as the contraction: per qubit one forward at batch S x G"""
    n, S, G = len(states), u.shape[0], len(grid)
    ident = torch.eye(K, dtype=torch.float64).expand(S * G, K, K)
    mx_grid = oc.generate_data(torch.tensor(grid, dtype=torch.float64).unsqueeze(1), K, dtype=torch.float64)[0][0]   # (G,K,K)
    chosen = [None] * n
    out = np.zeros((S, n))
    for q in range(n):
        mats = []
        for i in range(n):
            if i == q:
                mats.append(mx_grid.unsqueeze(0).expand(S, G, K, K).reshape(S * G, K, K))
            elif i < q:
                mats.append(chosen[i].unsqueeze(1).expand(S, G, K, K).reshape(S * G, K, K))
            else:
                mats.append(ident)
        res = oc.forward(graph, cores, states, mats).reshape(S, G).numpy()
        dens = np.maximum(res, 0.0)
        for s in range(S):
            cdf = np.cumsum(dens[s])
            cdf = cdf / (cdf[-1] + 1e-10)
            out[s, q] = _interp(cdf, u[s, q], grid)
        chosen[q] = oc.generate_data(torch.tensor(out[:, q:q + 1]), K, dtype=torch.float64)[0][0]
    return out


@pytest.mark.parametrize("n,K,G,S", [(4, 3, 24, 6), (3, 2, 17, 5), (5, 3, 12, 4), (2, 4, 15, 4)])
def test_prefix_environment_sampling_equals_the_reference_procedure(n, K, G, S):
    graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
    names, table, nq = oc.parse_graph(graph)
    torch.manual_seed(n * 10 + K)
    cores = {k: v.double() for k, v in oc.random_cores(table).items()}
    states = [s.double() for s in oc.unit_states(nq, K)]
    rng = np.random.default_rng(5)
    u = rng.uniform(0.02, 0.98, size=(S, n))
    grid = np.linspace(-3.0, 3.0, G)
    # the chain's cores in wire order: the core whose first qubit is q (the order csrc/tnq_chain.cu receives them in)
    q_mirror = tneq_b200.QCTN(graph)
    first = {e["core_name"]: min(x["qubit_idx"] for x in e["in_edge_list"]) for e in q_mirror.adjacency_table}
    order = sorted(q_mirror.cores, key=lambda c: first[c])
    chain = [cores[c].numpy() for c in order]
    got = prefix_sample(chain, [s.numpy() for s in states], u, grid, K)
    want = reference_sample(graph, cores, states, u, grid, K)
    assert np.abs(got - want).max() < 1e-9, np.abs(got - want).max()
