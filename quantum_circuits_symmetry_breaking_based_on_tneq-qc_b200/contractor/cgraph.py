"""Contraction graph: the greedy schedule lowered to PAIRWISE real contractions.

Input : a GreedySchedule (one multi-operand einsum per qubit group, exactly the
        reference's bookkeeping, greedy_strategy.py:690-990).
Output: a DAG of nodes over REAL tensors
          input     a tensor handed in by the caller (core, state, Mx, grad seed)
          contract  C = P o Q  (sum over the indices P and Q share)
          lin       sparse linear map with <= 2 terms per destination element
                    (complex conjugate, complex->2x2-real expansion, permute,
                    and the adjoints of those)
          seed      d(loss)/d(result) from the fused loss
        plus, if gradients are requested, the reverse-mode adjoint nodes.

Design points (DESIGN.md section "plan compiler"):
  * batch symbols 'a'/'b' are not indices here: a node is either `batched`
    (one value per sample) or shared.  Shared sub-trees (cores contracted with
    circuit states, complex expansion of cores, ...) are batch independent and
    end up in the one-off PREP section of the device program.
  * complex tensors are real tensors with a trailing index of extent 2.
    P o Q over complex numbers becomes ONE real contraction of P with the
    2x2-real expansion of Q:  Qx[k,ri,n,ro] = E[ri,ro,c] Q[k,n,c]; so the device
    only ever runs real arithmetic, and reverse mode in this representation
    directly yields PyTorch's complex gradient convention dL/dRe + i dL/dIm.
  * within a group the pairwise order is chosen by dynamic programming over
    operand subsets minimising per-sample multiply-adds (shared o shared
    products are nearly free); circuit states are folded into their core first.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .greedy_plan import GreedySchedule, Operand


@dataclass
class Node:
    id: int
    kind: str                       # input | contract | lin | seed
    idx: Tuple[int, ...]            # index ids in memory order (compact row-major)
    batched: bool
    cplx: bool = False              # last index is the (re, im) pair of a complex tensor
    needs_grad: bool = False
    operand: Optional[Operand] = None       # input
    p: int = -1                     # contract / lin source
    q: int = -1
    reduce_batch: bool = False      # contract of two batched tensors summed over samples
    lin_kind: str = ""              # lin: permute | conj | expand | fold | table  (tables are built lazily:
    lin_args: tuple = ()            #      large-bond tensors never materialise them, see CGraph.lin_tables)
    src_flat: Optional[np.ndarray] = None   # lin 'table': [ndst, T] flat source positions
    coef: Optional[np.ndarray] = None       # lin 'table': [ndst, T]
    acc_into: int = -1              # adjoint contribution accumulated into that node's buffer
    is_accum: bool = False          # zero-initialised accumulation buffer (shared adjoints)
    role: str = "fwd"               # fwd | adj
    tag: str = ""


class CGraph:
    def __init__(self, complex_mode: bool):
        self.complex_mode = complex_mode
        self.dims: Dict[int, int] = {}
        self.nodes: List[Node] = []
        self.inputs: Dict[Tuple[str, object], List[int]] = {}
        self.result: int = -1
        self.result_batch: str = ""
        self.seed: int = -1
        self.grads: Dict[Tuple[str, object], int] = {}     # input key -> accum node (GOUT)
        self.flops_per_sample = 0.0
        self.flops_shared = 0.0

    # ---- construction helpers ---------------------------------------------
    def new_index(self, extent: int) -> int:
        i = len(self.dims)
        self.dims[i] = int(extent)
        return i

    def size(self, idx: Sequence[int]) -> int:
        n = 1
        for i in idx:
            n *= self.dims[i]
        return n

    def shape(self, idx):
        return tuple(self.dims[i] for i in idx)

    def _add(self, **kw) -> Node:
        n = Node(id=len(self.nodes), **kw)
        self.nodes.append(n)
        return n

    def add_input(self, operand: Operand, idx, batched, needs_grad=False, cplx=False) -> Node:
        n = self._add(kind="input", idx=tuple(idx), batched=batched, cplx=cplx, needs_grad=needs_grad,
                      operand=operand)
        return n

    def add_contract(self, p: Node, q: Node, out_idx=None, reduce_batch=False, role="fwd", cplx=False) -> Node:
        shared = [i for i in p.idx if i in q.idx]
        if out_idx is None:
            out_idx = [i for i in p.idx if i not in shared] + [i for i in q.idx if i not in shared]
        batched = (p.batched or q.batched) and not reduce_batch
        n = self._add(kind="contract", idx=tuple(out_idx), batched=batched, cplx=cplx,
                      needs_grad=p.needs_grad or q.needs_grad, p=p.id, q=q.id,
                      reduce_batch=reduce_batch, role=role)
        work = float(self.size(set(p.idx) | set(q.idx)))
        if p.batched or q.batched:
            self.flops_per_sample += 2.0 * work
        else:
            self.flops_shared += 2.0 * work
        return n

    def add_lin(self, src: Node, dst_idx, kind: str, args=(), role="fwd", cplx=False, batched=None,
                src_flat=None, coef=None) -> Node:
        return self._add(kind="lin", idx=tuple(dst_idx), batched=src.batched if batched is None else batched,
                         cplx=cplx, needs_grad=src.needs_grad, p=src.id, lin_kind=kind, lin_args=tuple(args),
                         src_flat=src_flat, coef=coef, role=role)

    # ---- elementary linear maps ---------------------------------------------
    def _flat_of(self, idx_from, idx_to):
        """For every element of a tensor laid out with index order `idx_to`, the flat
        position of the same element in layout `idx_from` (same index set)."""
        shp = self.shape(idx_to)
        grids = np.indices(shp).reshape(len(shp), -1) if shp else np.zeros((0, 1), dtype=np.int64)
        pos = {i: k for k, i in enumerate(idx_to)}
        flat = np.zeros(grids.shape[1] if shp else 1, dtype=np.int64)
        stride = 1
        for i in reversed(idx_from):
            flat = flat + grids[pos[i]] * stride
            stride *= self.dims[i]
        return flat

    def lin_permute(self, src: Node, dst_idx, role="fwd") -> Node:
        return self.add_lin(src, dst_idx, "permute", role=role, cplx=src.cplx)

    def lin_conj(self, src: Node, role="fwd") -> Node:
        assert src.cplx
        return self.add_lin(src, src.idx, "conj", role=role, cplx=True)

    def lin_expand(self, src: Node, ri_in: int, ro: int) -> Node:
        """Qx[x..., ri, ro] = E[ri, ro, c] Q[x..., c] with E the 2x2 real form of a
        complex scalar: [[re, im], [-im, re]] (rows ri, columns ro)."""
        assert src.cplx
        return self.add_lin(src, tuple(src.idx[:-1]) + (ri_in, ro), "expand", cplx=False)

    def lin_fold(self, g_expanded: Node, like: Node, role="adj") -> Node:
        """Adjoint of lin_expand: dQ[x..., c] from dQx[x..., ri, ro]."""
        return self.add_lin(g_expanded, like.idx, "fold", role=role, cplx=True)

    def lin_tables(self, n: Node):
        """(src_flat [ndst, T], coef [ndst, T]) of a lin node -- only the shared-memory VM path
        asks for them; tensors are small there."""
        src = self.nodes[n.p]
        nd = self.size(n.idx)
        if n.lin_kind == "table":
            return n.src_flat, n.coef
        if n.lin_kind == "permute":
            flat = self._flat_of(src.idx, n.idx)
            return flat[:, None], np.ones((nd, 1))
        if n.lin_kind == "conj":
            flat = np.arange(nd, dtype=np.int64)
            return flat[:, None], np.where(flat % 2 == 0, 1.0, -1.0)[:, None]
        if n.lin_kind == "expand":
            nb = self.size(src.idx[:-1])
            base = np.arange(nb, dtype=np.int64)[:, None] * 2
            c_of = np.array([0, 1, 1, 0], dtype=np.int64)[None, :]          # (ri, ro) -> c
            sg = np.array([1.0, 1.0, -1.0, 1.0])[None, :]
            return (base + c_of).reshape(-1)[:, None], np.broadcast_to(sg, (nb, 4)).reshape(-1)[:, None]
        if n.lin_kind == "fold":
            nb = nd // 2
            base = np.arange(nb, dtype=np.int64)[:, None] * 4                 # position of (x, ri=0, ro=0)
            # c = 0: (0,0) + (1,1) ; c = 1: (0,1) - (1,0)
            s0 = (base + np.array([0, 1])[None, :]).reshape(-1)
            s1 = (base + np.array([3, 2])[None, :]).reshape(-1)
            c0 = np.ones(nd)
            c1 = np.tile(np.array([1.0, -1.0]), nb)
            return np.stack([s0, s1], axis=1), np.stack([c0, c1], axis=1)
        raise ValueError(n.lin_kind)

    # ---- complex-aware contraction -------------------------------------------
    def contract(self, p: Node, q: Node) -> Node:
        if not self.complex_mode:
            return self.add_contract(p, q)
        assert p.cplx and q.cplx
        # expand the operand that plays the "matrix" role: shared beats batched, small beats big
        def weight(n):
            return (1 if n.batched else 0, self.size(n.idx))
        x, y = (q, p) if weight(q) <= weight(p) else (p, q)
        ro = self.new_index(2)
        xx = self.lin_expand(x, y.idx[-1], ro)
        shared = [i for i in y.idx if i in xx.idx]
        out = [i for i in y.idx if i not in shared] + [i for i in xx.idx if i not in shared and i != ro] + [ro]
        return self.add_contract(y, xx, out_idx=out, cplx=True)


# ----------------------------------------------------------------------------
# pairwise order inside one greedy group
# ----------------------------------------------------------------------------
def _pairwise_order(index_sets, batched, out_set, dims):
    """Return a list of (i, j) merges over a growing operand list (opt_einsum
    'path' convention is avoided: positions refer to an append-only list).
    Minimises sum of per-contraction multiply-adds; shared o shared costs 1e-6x."""
    n = len(index_sets)
    if n == 1:
        return []

    def extent(s):
        v = 1.0
        for i in s:
            v *= dims[i]
        return v

    if n <= 12:
        full = (1 << n) - 1
        cover = {}
        for m in range(1, full + 1):
            inside = set()
            for k in range(n):
                if m >> k & 1:
                    inside |= index_sets[k]
            cover[m] = inside
        res, isb = {}, {}
        for m in range(1, full + 1):
            outside = set(out_set)
            for k in range(n):
                if not m >> k & 1:
                    outside |= index_sets[k]
            res[m] = frozenset(i for i in cover[m] if i in outside)
            isb[m] = any(batched[k] for k in range(n) if m >> k & 1)
        best = {1 << k: (0.0, None) for k in range(n)}
        for size in range(2, n + 1):
            for combo in itertools.combinations(range(n), size):
                m = 0
                for k in combo:
                    m |= 1 << k
                low = m & -m
                sub = (m - 1) & m
                cand = None
                while sub:
                    if sub & low:
                        rest = m ^ sub
                        if rest and sub in best and rest in best:
                            a, b = res[sub], res[rest]
                            connected = bool(a & b)
                            cost = extent(a | b) * (1.0 if (isb[sub] or isb[rest]) else 1e-6)
                            if not connected:
                                cost *= 1e3  # outer products only as a last resort
                            tot = best[sub][0] + best[rest][0] + cost
                            if cand is None or tot < cand[0]:
                                cand = (tot, (sub, rest))
                    sub = (sub - 1) & m
                if cand is not None:
                    best[m] = cand
        merges, slot = [], {1 << k: k for k in range(n)}

        def emit(m):
            if m in slot:
                return slot[m]
            a, b = best[m][1]
            ia, ib = emit(a), emit(b)
            slot[m] = n + len(merges)
            merges.append((ia, ib))
            return slot[m]

        emit(full)
        return merges
    # greedy fallback for very large groups
    live = {k: (frozenset(index_sets[k]), batched[k]) for k in range(n)}
    merges, nxt = [], n
    while len(live) > 1:
        pick = None
        keys = list(live)
        for x in range(len(keys)):
            for y in range(x + 1, len(keys)):
                a, b = live[keys[x]], live[keys[y]]
                if not (a[0] & b[0]) and len(live) > 2:
                    continue
                c = extent(a[0] | b[0]) * (1.0 if (a[1] or b[1]) else 1e-6)
                if pick is None or c < pick[0]:
                    pick = (c, keys[x], keys[y])
        if pick is None:
            pick = (0.0, keys[0], keys[1])
        _, kx, ky = pick
        a, b = live.pop(kx), live.pop(ky)
        others = set(out_set)
        for v in live.values():
            others |= v[0]
        live[nxt] = (frozenset(i for i in (a[0] | b[0]) if i in others), a[1] or b[1])
        merges.append((kx, ky))
        nxt += 1
    return merges


def build_forward(schedule: GreedySchedule, complex_mode: bool, core_shapes: Dict[object, Tuple[int, ...]],
                  trainable: Sequence[Tuple[str, object]] = ()) -> CGraph:
    """Lower the schedule.  core_shapes: ('core'|'rcore', name) -> dims.
    `trainable`: input keys ('core', name) / ('rcore', name) that require gradients.

    Index ids are GLOBAL: one id per network edge (the schedule's raw symbols), so
    a tensor produced at one qubit is consumed later without any relabelling.  A
    core that occurs as its L copy and as its R copy is two input nodes over the
    same caller buffer (their gradients accumulate into one output buffer)."""
    g = CGraph(complex_mode)
    trainable = set(trainable)
    edge_id: Dict[str, int] = {}
    produced: Dict[int, Node] = {}
    want_order: Dict[int, List[int]] = {}

    def ids_of(symbols, dims):
        out = []
        for c in symbols:
            if c not in edge_id:
                edge_id[c] = g.new_index(dims[c])
            out.append(edge_id[c])
        return out

    for step in schedule.steps:
        nodes, state_pos = [], []
        for sub, op in zip(step.raw_subs, step.operands):
            body = [c for c in sub if c not in "ab"]
            ids = ids_of(body, step.raw_dims)
            if op.kind == "tmp":
                n = produced[op.key]
                have = n.idx[:-1] if n.cplx else n.idx
                assert sorted(have) == sorted(ids), "tmp operand does not carry the expected edges"
            else:
                key = ("core", op.key) if op.kind in ("core", "core_conj") else (op.kind, op.key)
                if key[0] in ("core", "rcore"):
                    if tuple(core_shapes[key]) != tuple(g.dims[i] for i in ids):
                        raise ValueError(f"core {op.key!r}: shape {tuple(core_shapes[key])} does not match its "
                                         f"edges {tuple(g.dims[i] for i in ids)}")
                idx = list(ids) + ([g.new_index(2)] if complex_mode else [])
                n = g.add_input(Operand(key[0], key[1]), idx, batched=any(c in "ab" for c in sub),
                                needs_grad=key in trainable, cplx=complex_mode)
                g.inputs.setdefault(key, []).append(n.id)
                if op.kind == "core_conj" and complex_mode:
                    n = g.lin_conj(n)
                if op.kind == "state":
                    state_pos.append(len(nodes))
            nodes.append(n)
        out_ids = ids_of([c for c in step.raw_out if c not in "ab"], step.raw_dims)
        # fold circuit states into the tensor that carries their edge
        alive = {k: n for k, n in enumerate(nodes) if k not in state_pos}
        for k in state_pos:
            edge = nodes[k].idx[0]
            host = next((h for h, n in alive.items() if edge in n.idx), None)
            if host is None:
                alive[k] = nodes[k]
            else:
                alive[host] = g.contract(alive[host], nodes[k])
        live = [alive[k] for k in sorted(alive)]
        sets = [set(n.idx[:-1] if n.cplx else n.idx) for n in live]
        # An edge that occurs in ONE operand and not in the output is summed over by einsum without a
        # partner (the reference's wiring of `right_qctn=<QCTN>` leaves such edges,
        # greedy_strategy.py:764-822): contract it with a vector of ones -- an input like a circuit state.
        for i in sorted(set().union(*sets)):
            if sum(i in s for s in sets) == 1 and i not in out_ids:
                k = next(k for k, s in enumerate(sets) if i in s)
                okey = (len(g.inputs), i)
                ones = g.add_input(Operand("ones", okey), [i] + ([g.new_index(2)] if complex_mode else []),
                                   batched=False, cplx=complex_mode)
                g.inputs[("ones", okey)] = [ones.id]
                live[k] = g.contract(live[k], ones)
                sets[k] = set(live[k].idx[:-1] if live[k].cplx else live[k].idx)
        pool = list(live)
        for a, b in _pairwise_order(sets, [n.batched for n in live], set(out_ids), g.dims):
            pool.append(g.contract(pool[a], pool[b]))
        res = pool[-1]
        res.tag = f"step{step.out}"
        produced[step.out] = res
        want_order[step.out] = out_ids
    if schedule.result.kind != "tmp":
        raise RuntimeError("nothing to contract: the network is a single tensor")
    res = produced[schedule.result.key]
    want = want_order[schedule.result.key]
    if list(res.idx[:-1] if res.cplx else res.idx) != want:
        res = g.lin_permute(res, tuple(want) + ((res.idx[-1],) if res.cplx else ()))
    if res.kind == "input":
        res = g.lin_permute(res, res.idx)      # single-operand network: plain copy
    g.result = res.id
    g.result_batch = "".join(c for c in "ab" if any(c in sub for st in schedule.steps for sub in st.raw_subs))
    return g


# ----------------------------------------------------------------------------
# reverse mode
# ----------------------------------------------------------------------------
def add_backward(g: CGraph, seed_from: str):
    """Append adjoint nodes.  seed_from: 'input' (d result supplied by the caller,
    autograd route) or 'loss' (fused clamp/log/mean loss, engine_siamese.py:490-530).

    Batched tensors have exactly one consumer, so their adjoint is one node.
    Shared tensors (cores and everything derived from them in PREP) may have
    several; their adjoints are zero-initialised accumulation buffers that every
    contribution adds into.  All occurrences of one caller tensor share ONE
    accumulation buffer (g.grads[key]): that buffer is the returned gradient."""
    res = g.nodes[g.result]
    if seed_from == "input":
        seed = g.add_input(Operand("gradseed", 0), res.idx, True, cplx=res.cplx)
        g.inputs[("gradseed", 0)] = [seed.id]
    else:
        seed = g._add(kind="seed", idx=res.idx, batched=True, cplx=res.cplx, p=res.id, role="adj")
    g.seed = seed.id
    adj: Dict[int, Node] = {res.id: seed}
    accum: Dict[object, Node] = {}

    def accum_for(target: Node) -> Node:
        key = ("in",) + (target.operand.kind, target.operand.key) if target.kind == "input" else ("node", target.id)
        if key not in accum:
            accum[key] = g._add(kind="lin", idx=target.idx, batched=False, cplx=target.cplx, is_accum=True,
                                role="adj", tag=f"accum:{key}")
            if target.kind == "input":
                g.grads[(target.operand.kind, target.operand.key)] = accum[key].id
        return accum[key]

    def adjoint_of(node: Node) -> Optional[Node]:
        if node.batched:
            return adj.get(node.id)
        key = ("node", node.id)
        return accum.get(key)

    def contribute(target: Node, make):
        if target.batched:
            assert target.id not in adj, "a batched tensor with two consumers is not supported"
            adj[target.id] = make(-1)
        else:
            make(accum_for(target).id)

    n_fwd = len(g.nodes)
    for node in reversed(g.nodes[:n_fwd]):
        if not node.needs_grad or node.kind in ("input", "seed") or node.role != "fwd":
            continue
        gnode = adjoint_of(node)
        if gnode is None:
            continue
        if node.kind == "contract":
            p, q = g.nodes[node.p], g.nodes[node.q]
            for me, other in ((p, q), (q, p)):
                if not me.needs_grad:
                    continue

                def make(acc, me=me, other=other):
                    red = (not me.batched) and (gnode.batched or other.batched)
                    c = g.add_contract(gnode, other, out_idx=me.idx, reduce_batch=red, role="adj", cplx=me.cplx)
                    c.acc_into, c.needs_grad = acc, False
                    return c

                contribute(me, make)
        elif node.kind == "lin":
            src = g.nodes[node.p]

            def make(acc, src=src, node=node):
                if node.lin_kind == "permute":
                    c = g.lin_permute(gnode, src.idx, role="adj")
                elif node.lin_kind == "conj":
                    c = g.lin_conj(gnode, role="adj")
                elif node.lin_kind == "expand":
                    c = g.lin_fold(gnode, src)
                else:
                    raise NotImplementedError(f"adjoint of lin '{node.lin_kind}'")
                c.acc_into, c.needs_grad = acc, False
                return c

            contribute(src, make)
    for key, ids in g.inputs.items():
        if g.nodes[ids[0]].needs_grad and key not in g.grads:
            raise RuntimeError(f"no gradient path to {key}")
    return g
