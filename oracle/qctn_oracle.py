"""CPU oracle for the QCTN contraction hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement, on the CPU, of what the reference
(`tneq_qc`, pure Python + torch) computes on the one path this repository
accelerates.  It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The product
package never does (it fails loudly when its CUDA library is missing).

Parity pin (see DESIGN.md "Oracle"): `oracle/make_golden.py` imports the real
reference from /root/reference in the build container and checks that this
restatement is BIT-IDENTICAL to it on CPU (same einsum strings, same operand
order, same torch kernels), then commits small fixtures under tests/golden/.

What is restated, with the reference file:line each piece follows:

  parse_graph            tneq_qc/core/qctn.py:482-536, 591-714
  TNT (TNTensor)         tneq_qc/core/tn_tensor.py:4-125
  greedy_contract        tneq_qc/contractor/greedy_strategy.py:41-600
      _components        ...greedy_strategy.py:615-664
      _fetch             ...greedy_strategy.py:667-687
      _contract_group    ...greedy_strategy.py:690-990
      _contract_rest     ...greedy_strategy.py:993-1080
  forward                tneq_qc/core/engine_siamese.py:261-349
  loss_and_grads         tneq_qc/core/engine_siamese.py:351-554 and
                         tneq_qc/backends/backend_pytorch.py:107-166
  generate_data          tneq_qc/core/engine_siamese.py:59-254
  init_random_core       tneq_qc/backends/backend_pytorch.py:470-495
  sgdg_step              tneq_qc/backends/backend_pytorch.py:200-268, 349-468

Arithmetic lives in torch (CPU): `torch.einsum` with opt_einsum NOT visible to
torch, i.e. operands contracted strictly left to right.
"""

from __future__ import annotations

import math
import random
import re
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

_BASE = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def get_symbol(i: int) -> str:
    """opt_einsum.parser.get_symbol (published behaviour, opt_einsum 3.x)."""
    if i < 52:
        return _BASE[i]
    if i >= 55296:
        return chr(i + 2048)
    return chr(i + 140)


# --------------------------------------------------------------------------
# TNTensor (tn_tensor.py:4-125)
# --------------------------------------------------------------------------
class TNT:
    """tensor * scale with the scale (and its log) carried on the host."""

    def __init__(self, tensor, scale=1.0, log_scale=None):
        self.tensor = tensor
        self.scale = float(scale)
        if log_scale is None:
            log_scale = math.log(abs(self.scale)) if self.scale != 0 else float("-inf")
        self.log_scale = log_scale

    ndim = property(lambda s: s.tensor.ndim)
    shape = property(lambda s: s.tensor.shape)
    dtype = property(lambda s: s.tensor.dtype)

    def auto_scale(self):
        m = self.tensor.abs().max()
        m = m.item() if hasattr(m, "item") else float(m)
        if m == 0:
            return
        self.tensor /= m
        self.scale *= m
        self.log_scale += math.log(abs(m))

    def scale_to(self, new_scale):
        new_scale = float(new_scale)
        self.tensor = self.tensor * (self.scale / new_scale)
        self.scale = new_scale
        self.log_scale = math.log(abs(new_scale))


def _raw(t):
    return t.tensor if isinstance(t, TNT) else t


# --------------------------------------------------------------------------
# graph string -> adjacency table (qctn.py:482-536, 591-714)
# --------------------------------------------------------------------------
def parse_graph(graph: str):
    """Return (core_names, table). table[i] = dict(name, ins, outs) where each
    edge is dict(nbr, rank, qubit); nbr == -1 marks a circuit input/output."""
    lines = graph.strip().splitlines()
    order = {get_symbol(i): i for i in range(10000)}
    names = sorted({ch for ch in graph if ch in order}, key=order.__getitem__)
    pos = {c: i for i, c in enumerate(names)}
    table = [dict(name=c, ins=[], outs=[]) for c in names]
    cls = "".join(names)
    first = re.compile(rf"^(\d+)([{cls}])")
    last = re.compile(rf"([{cls}])(\d+)$")
    link = re.compile(rf"([{cls}])(\d+)(?=[{cls}])")
    for q, line in enumerate(lines):
        s = line.strip().replace("-", "")
        r_in, c_in = first.match(s).groups()
        c_out, r_out = last.search(s).groups()
        table[pos[c_in]]["ins"].append(dict(nbr=-1, rank=int(r_in), qubit=q))
        table[pos[c_out]]["outs"].append(dict(nbr=-1, rank=int(r_out), qubit=q))
        for m in link.finditer(s):
            if m.end() >= len(s):
                break
            src, rank = m.groups()
            dst = s[m.end()]
            table[pos[src]]["outs"].append(dict(nbr=pos[dst], rank=int(rank), qubit=q))
            table[pos[dst]]["ins"].append(dict(nbr=pos[src], rank=int(rank), qubit=q))
    return names, table, len(lines)


def core_shapes(table):
    return {t["name"]: [e["rank"] for e in t["ins"]] + [e["rank"] for e in t["outs"]] for t in table}


def init_random_core(shape, dtype=torch.float32):
    """QR-orthogonal init (backend_pytorch.py:470-495)."""
    d = int(np.prod(shape[: len(shape) // 2]))
    q, r = torch.linalg.qr(torch.randn((d, d), dtype=dtype))
    diag = torch.diag(r)
    if torch.is_complex(diag):
        q = q @ torch.diag((diag / (diag.abs() + 1e-12)).conj())
    else:
        q = q * torch.sign(diag).unsqueeze(0)
    return q.reshape(shape)


def random_cores(table, dtype=torch.float32):
    out = {}
    for t in table:
        din = int(np.prod([e["rank"] for e in t["ins"]])) if t["ins"] else 1
        dout = int(np.prod([e["rank"] for e in t["outs"]])) if t["outs"] else 1
        out[t["name"]] = init_random_core([din, dout], dtype).reshape(core_shapes([t])[t["name"]])
    return out


# --------------------------------------------------------------------------
# greedy qubit-by-qubit contraction (greedy_strategy.py)
# --------------------------------------------------------------------------
class _N:
    """One tensor of the symmetric L / M / R network."""

    __slots__ = ("uid", "name", "src", "key", "ins", "outs", "side", "batch", "tensor")

    def __init__(self, uid, name, src, key, ins, outs, side, batch=""):
        self.uid, self.name, self.src, self.key = uid, name, src, key
        self.ins, self.outs, self.side, self.batch = ins, outs, side, batch
        self.tensor = None

    def edges(self):
        return self.ins + self.outs


def _cp(edges):
    return [dict(e) for e in edges]


def _present(container, q):
    if container is None:
        return False
    if isinstance(container, dict):
        return q in container
    if isinstance(container, (list, tuple)):
        return q < len(container)
    return True


def _fetch(node, cores, states, mxs):
    if node.tensor is not None:
        return node.tensor
    if node.src == "core":
        return cores[node.key]
    if node.src == "transpose":
        c = cores[node.key]
        return c.conj() if _raw(c).is_complex() else c  # D2 of SURVEY: needs TNT.conj
    if node.src == "circuit":
        return states[node.key]
    if node.src == "mx":
        return mxs[node.key]
    raise ValueError(node.src)


def _tnt_conj(self):
    return TNT(self.tensor.conj(), self.scale, self.log_scale)


TNT.conj = _tnt_conj
TNT.is_complex = lambda self: self.tensor.is_complex()


def _components(members):
    """Connected components in first-appearance order (greedy_strategy.py:615-664)."""
    n = len(members)
    if n <= 1:
        return [members] if n else []
    where = {m.uid: i for i, m in enumerate(members)}
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i, m in enumerate(members):
        for e in m.outs + m.ins:
            j = where.get(e["nbr"]) if e["nbr"] >= 0 else None
            if j is not None:
                a, b = find(i), find(j)
                if a != b:
                    parent[a] = b
    groups: Dict[int, list] = {}
    for i in range(n):
        groups.setdefault(find(i), []).append(members[i])
    return list(groups.values())


def _remap(eq: str) -> str:
    table = {"a": "a", "b": "b", ",": ",", "-": "-", ">": ">"}
    nxt = 2
    for ch in eq:
        if ch not in table:
            table[ch] = get_symbol(nxt)
            nxt += 1
    return "".join(table[ch] for ch in eq)


def _run_einsum(eq, tensors, log):
    raws, scale, lscale, wrapped = [], None, None, False
    for t in tensors:
        if isinstance(t, TNT):
            wrapped = True
            raws.append(t.tensor)
            scale = t.scale if scale is None else scale * t.scale
            lscale = t.log_scale if lscale is None else lscale + t.log_scale
        else:
            raws.append(t)
    if log is not None:
        log.append((eq, [tuple(r.shape) for r in raws]))
    out = torch.einsum(eq, *raws)
    return TNT(out, scale, lscale) if wrapped else out


def _contract_group(group, q, cores, states, mxs, log):
    if len(group) == 1 and not any(e["qubit"] == q for e in group[0].edges()):
        return group[0]
    inside = {m.uid for m in group}
    parts, tensors, keep_in, keep_out, batch = [], [], [], [], set()

    def visit(e, bucket):
        internal = e["nbr"] >= 0 and e["nbr"] in inside
        if e["nbr"] == -1 or (not internal and e["qubit"] != q):
            bucket.append(dict(e))

    for m in group:
        t = _fetch(m, cores, states, mxs)
        tensors.append(t)
        batch.update(m.batch)
        if m.side == "R":
            n_in, n_out = len(m.outs), len(m.ins)  # counts of the ORIGINAL core
            dims = [None] * (t.ndim - len(m.batch))
            for i, e in enumerate(m.outs):
                dims[n_in - 1 - i] = e["sym"]
                visit(e, keep_out)
            for i, e in enumerate(m.ins):
                dims[n_in + n_out - 1 - i] = e["sym"]
                visit(e, keep_in)
            parts.append(m.batch + "".join(s for s in dims if s is not None))
        else:
            part = m.batch
            for e in m.ins:
                part += e["sym"]
                visit(e, keep_in)
            for e in m.outs:
                part += e["sym"]
                visit(e, keep_out)
            parts.append(part)
    new_batch = "".join(c for c in "ab" if c in batch)
    out = new_batch + "".join(e["sym"] for e in keep_in) + "".join(e["sym"] for e in keep_out)
    eq = _remap(",".join(parts) + "->" + out)
    merged = _N(-1 - q, f"merged_{q}", "merged", None, keep_in, keep_out, "M", new_batch)
    merged.tensor = _run_einsum(eq, tensors, log)
    return merged


def _contract_rest(nodes, cores, states, mxs, log):
    tensors = [_fetch(n, cores, states, mxs) for n in nodes]
    parts, outsyms = [], []
    for n, t in zip(nodes, tensors):
        if n.side == "M":
            nb = t.ndim - 2
            part = ("a" if nb >= 1 else "") + ("b" if nb >= 2 else "")
            if n.ins:
                part += n.ins[0]["sym"]
            if n.outs:
                part += n.outs[0]["sym"]
        elif n.side == "R":
            n_in, n_out = len(n.outs), len(n.ins)
            dims = [None] * t.ndim
            for i, e in enumerate(n.outs):
                dims[n_in - 1 - i] = e["sym"]
            for i, e in enumerate(n.ins):
                dims[n_in + n_out - 1 - i] = e["sym"]
            part = "".join(s for s in dims if s is not None)
        else:
            part = "".join(e["sym"] for e in n.ins) + "".join(e["sym"] for e in n.outs)
            if n.src == "merged":
                extra = t.ndim - len(n.ins) - len(n.outs)
                part = ("a" if extra >= 1 else "") + ("b" if extra >= 2 else "") + part
        parts.append(part)
        for c in "ab":
            if c in part and c not in outsyms:
                outsyms.append(c)
    eq = ",".join(parts) + "->" + "".join(outsyms)
    if log is not None:
        log.append((eq, [tuple(_raw(t).shape) for t in tensors]))
    return torch.einsum(eq, *tensors)


def greedy_contract(table, nqubits, cores, states, mxs, right="symmetric", right_table=None,
                    right_cores=None, log: Optional[list] = None):
    """Build the L / M / R network and contract it one qubit at a time."""
    cores = dict(cores)
    nodes: List[_N] = []
    lmap, smapL, mmap, rmap, smapR = {}, {}, {}, {}, {}
    for i, t in enumerate(table):
        lmap[i] = len(nodes)
        nodes.append(_N(len(nodes), t["name"] + "_L", "core", t["name"], _cp(t["ins"]), _cp(t["outs"]), "L"))
    for q in range(nqubits):
        if _present(states, q):
            smapL[q] = len(nodes)
            e = dict(nbr=-1, rank=states[q].shape[0], qubit=q)
            nodes.append(_N(len(nodes), f"circuit_L_{q}", "circuit", q, [], [e], "L"))
    for q in range(nqubits):
        if _present(mxs, q) and mxs[q] is not None:
            mx = mxs[q]
            mmap[q] = len(nodes)
            b = "a" if mx.ndim == 3 else ("ab" if mx.ndim == 4 else "")
            nodes.append(_N(len(nodes), f"mx_{q}", "mx", q,
                            [dict(nbr=-1, rank=mx.shape[-2], qubit=q)],
                            [dict(nbr=-1, rank=mx.shape[-1], qubit=q)], "M", b))
    if isinstance(right, str) and right == "symmetric":
        for i, t in enumerate(table):
            rmap[i] = len(nodes)
            nodes.append(_N(len(nodes), t["name"] + "_R", "transpose", t["name"],
                            _cp(t["outs"])[::-1], _cp(t["ins"])[::-1], "R"))
    elif right == "qctn":
        for i, t in enumerate(right_table):
            cores["right_" + t["name"]] = right_cores[t["name"]]
            rmap[i + len(lmap)] = len(nodes)
            nodes.append(_N(len(nodes), t["name"] + "_R", "core", "right_" + t["name"],
                            _cp(t["ins"]), _cp(t["outs"]), "R"))
    elif right is not None:
        raise ValueError("Invalid right_qctn parameter.")
    for q in range(nqubits):
        if _present(states, q):
            smapR[q] = len(nodes)
            e = dict(nbr=-1, rank=states[q].shape[0], qubit=q)
            nodes.append(_N(len(nodes), f"circuit_R_{q}", "circuit", q, [e], [], "R"))

    # wire neighbours (greedy_strategy.py:297-406)
    def wire(uid, edges, cmap, open_map, far_side):
        for e in edges:
            if e.get("is_cross_partition"):
                continue
            if e["nbr"] == -1:
                if e["qubit"] in open_map:
                    other = open_map[e["qubit"]]
                    e["nbr"] = other
                    getattr(nodes[other], far_side)[0]["nbr"] = uid
            elif e["nbr"] in cmap:
                e["nbr"] = cmap[e["nbr"]]

    for uid in lmap.values():
        wire(uid, nodes[uid].ins, lmap, smapL, "outs")
        wire(uid, nodes[uid].outs, lmap, mmap, "ins")
    for uid in rmap.values():
        wire(uid, nodes[uid].ins, rmap, mmap, "outs")
        wire(uid, nodes[uid].outs, rmap, smapR, "ins")

    # edge symbols, skipping the batch letters (greedy_strategy.py:408-449)
    def fresh():
        i = 0
        while True:
            s = get_symbol(i)
            if s not in ("a", "b"):
                yield s
            i += 1

    gen = fresh()
    for n in nodes:
        for e in n.outs:
            if "sym" in e:
                continue
            e["sym"] = next(gen)
            if e["nbr"] >= 0:
                for f in nodes[e["nbr"]].ins:
                    if f["nbr"] == n.uid and f["qubit"] == e["qubit"]:
                        f["sym"] = e["sym"]
                        break
    for n in nodes:
        for e in n.ins:
            if "sym" not in e:
                e["sym"] = next(gen)

    # qubit sweep (greedy_strategy.py:456-585)
    nxt = len(nodes)
    for q in range(nqubits):
        here = [n for n in nodes if any(e["qubit"] == q for e in n.edges())]
        if not here:
            continue
        extra = []
        for n in here:
            for e in n.ins + n.outs:
                if e["nbr"] < 0:
                    continue
                other = next((c for c in nodes if c.uid == e["nbr"]), None)
                if other is not None and other.src == "circuit" and \
                        not any(other is h for h in here) and not any(other is x for x in extra):
                    extra.append(other)
        here = here + extra
        made, gone, repl = [], set(), {}
        for gi, group in enumerate(_components(here)):
            new = _contract_group(group, q, cores, states, mxs, log)
            if new is None or any(new is m for m in group):
                continue
            new.uid, new.name = nxt, f"merged_q{q}_g{gi}_{nxt}"
            nxt += 1
            made.append(new)
            for m in group:
                gone.add(m.uid)
                repl[m.uid] = new
        if not made:
            continue
        nodes = [n for n in nodes if n.uid not in gone] + made
        for n in nodes:
            for e in n.edges():
                if e["nbr"] in repl:
                    e["nbr"] = repl[e["nbr"]].uid
    if not nodes:
        raise RuntimeError("No tensor left after contraction")
    if len(nodes) == 1:
        return _fetch(nodes[0], cores, states, mxs)
    return _contract_rest(nodes, cores, states, mxs, log)


# --------------------------------------------------------------------------
# engine level (engine_siamese.py)
# --------------------------------------------------------------------------
def abs_square(t):
    if torch.is_complex(t):
        return t.real * t.real + t.imag * t.imag
    return t


def forward(graph, cores, states, mxs, ret_type="tensor", log=None):
    """EngineSiamese.contract_with_compiled_strategy (engine_siamese.py:261-349).
    NOTE complex dtypes return |c|^2 of an already symmetric contraction
    (SURVEY D10); that quirk is part of the contract."""
    names, table, nq = parse_graph(graph)
    res = greedy_contract(table, nq, {n: cores[n] for n in names}, states, mxs, log=log)
    if isinstance(res, TNT):
        if ret_type == "TNTensor":
            if torch.is_complex(res.tensor):
                res = TNT(abs_square(res.tensor), res.scale, res.log_scale)
            return res
        res.scale_to(1.0)
        return abs_square(res.tensor)
    return abs_square(res)


def loss_and_grads(graph, cores, states, mxs, log=None):
    """EngineSiamese.contract_with_compiled_strategy_for_gradient
    (engine_siamese.py:351-554) driven by torch.autograd.grad
    (backend_pytorch.py:107-166).  Every core is treated as trainable.
    loss = -mean_b[ log(clamp(value_b, 1e-10)) + log_scale ]."""
    names, table, nq = parse_graph(graph)
    leaves, scales = [], []
    for n in names:
        c = cores[n]
        leaves.append(_raw(c).detach().clone().requires_grad_(True))
        scales.append(c.scale if isinstance(c, TNT) else 1.0)
    wrapped = {n: TNT(t, s) for n, t, s in zip(names, leaves, scales)}
    res = greedy_contract(table, nq, wrapped, states, mxs, log=log)
    val, lscale = (res.tensor, res.log_scale) if isinstance(res, TNT) else (res, 0.0)
    val = abs_square(val)
    target = torch.ones(val.shape, dtype=val.dtype)
    total = torch.log(torch.clamp(val, min=1e-10)) + lscale
    loss = -torch.mean(target * total)
    grads = torch.autograd.grad(loss, leaves)
    return loss.detach(), list(grads)


def hermite_weights(k_max):
    k = np.arange(k_max + 1, dtype=np.float64)
    lf = np.array([math.lgamma(int(i) + 1) for i in k], dtype=np.float64)
    return np.exp(-0.5 * (0.5 * math.log(2 * math.pi) + lf)).astype(np.float64)


def generate_data(x, K, dtype=torch.float32, ret_type="tensor", mx_K=None):
    """EngineSiamese.generate_data (engine_siamese.py:133-254).
    phi_k(x) = (2 pi)^-1/4 (k!)^-1/2 exp(-x^2/4) He_k(x); Mx = conj(phi) phi^T."""
    w_np = hermite_weights(max(K, mx_K or K))
    x = x.to(dtype)
    nq = x.shape[1]
    if dtype.is_complex:
        xr = np.asarray(x.detach().cpu().numpy().real, dtype=np.float64)
        H = np.zeros((K,) + xr.shape, dtype=np.float64)
        H[0] = 1.0
        if K >= 2:
            H[1] = xr
            for i in range(2, K):
                H[i] = xr * H[i - 1] - (i - 1) * H[i - 2]
        g = np.sqrt(np.exp(-np.square(xr) / 2.0))[..., None]
        phi = w_np[:K][None, None, :] * g * np.transpose(H, (1, 2, 0))
        M = np.einsum("bdk,bdl->bdkl", phi, phi)
        out = torch.as_tensor(phi, dtype=dtype)
        mats = [torch.as_tensor(M[:, i, :, :], dtype=dtype) for i in range(nq)]
    else:
        w = torch.as_tensor(w_np, dtype=dtype)[:K].unsqueeze(0).unsqueeze(0)
        H = torch.zeros((K,) + tuple(x.shape), dtype=x.dtype)
        H[0] = torch.ones_like(x)
        if K >= 2:
            H[1] = x
            for i in range(2, K):
                H[i] = x * H[i - 1] - (i - 1) * H[i - 2]
        g = torch.sqrt(torch.exp(-torch.square(x) / 2)).unsqueeze(-1)
        out = w * g * H.permute(1, 2, 0)
        M = torch.einsum("bdk,bdl->bdkl", out.conj(), out)
        mats = [M[:, i, :, :] for i in range(nq)]
    if ret_type == "TNTensor":
        wrapped = []
        for m in mats:
            t = TNT(m)
            t.auto_scale()
            wrapped.append(t)
        mats = wrapped
    return mats, out


def unit_states(nq, K, dtype=torch.float32):
    """e_{K-1} on every qubit (examples/example_train_single_node.py:46-54)."""
    out = [torch.zeros(K, dtype=dtype) for _ in range(nq)]
    for s in out:
        s[-1] = 1.0
    return out


# --------------------------------------------------------------------------
# SGDG / Cayley step (backend_pytorch.py:200-268, 349-468)  ["next" row]
# --------------------------------------------------------------------------
def sgdg_step(params, grads, state, lr=0.01, momentum=0.0, stiefel=True, rng=random):
    """One Stiefel-manifold SGD step on a list of plain tensors."""
    eps = 1e-8
    if "momentum_buffer" not in state:
        state["momentum_buffer"] = [None] * len(params)
    new = []
    with torch.no_grad():
        for i, (p, g) in enumerate(zip(params, grads)):
            shp = p.shape
            if len(shp) > 2:
                d = int(np.prod(shp[: len(shp) // 2]))
                p2, g2 = p.reshape(d, -1), g.reshape(d, -1)
            else:
                p2, g2 = p, g
            cplx = torch.is_complex(p2)
            unity = p2 / (torch.norm(p2, p=2, dim=1, keepdim=True) + eps)
            if stiefel and unity.shape[0] <= unity.shape[1]:
                if rng.randint(1, 101) == 1:
                    qq, rr = torch.linalg.qr(unity.T, mode="reduced")
                    dd = torch.diag(rr)
                    qq = qq * (torch.sgn(dd) if torch.is_complex(dd) else torch.sign(dd)).unsqueeze(0)
                    unity = qq.T
                if state["momentum_buffer"][i] is None:
                    state["momentum_buffer"][i] = torch.zeros(g2.T.shape, dtype=g2.dtype)
                V = state["momentum_buffer"][i]
                V = momentum * V - (torch.conj(g2).T if cplx else g2.T)
                MX = V @ unity
                XMX = unity @ MX
                uH = torch.conj(unity).T if cplx else unity.T
                W_hat = MX - 0.5 * (uH @ XMX)
                W = W_hat - (torch.conj(W_hat).T if cplx else W_hat.T)
                t = 0.5 * 2 / (torch.abs(W).sum(dim=0).max() + eps)
                alpha = min(t, lr)
                I = torch.eye(W.shape[0], dtype=W.dtype)
                Y = torch.inverse(I - (alpha / 2) * W) @ (I + (alpha / 2) * W) @ (uH if cplx else unity.T)
                pn = torch.conj(Y).T if cplx else Y.T
                new.append(pn.reshape(shp) if len(shp) > 2 else pn)
                state["momentum_buffer"][i] = W @ uH
            else:
                new.append(p - lr * g)
    return new, state
