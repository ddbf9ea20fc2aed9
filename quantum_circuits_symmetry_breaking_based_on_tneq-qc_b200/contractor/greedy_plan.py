"""Symbolic greedy schedule: WHICH tensors are contracted at WHICH qubit, with
WHICH einsum index string -- no tensors, no arithmetic.

The reference re-derives this bookkeeping with real tensors on every call
(tneq_qc/contractor/greedy_strategy.py:45-598, 3-9 ms of Python per forward).
Here it is derived once per (graph, operand signature) and cached; the device
plan is lowered from it.  The result must be IDENTICAL to the reference's
bookkeeping (same groups, same operand order, same remapped einsum strings):
tests/test_greedy_plan.py checks it against golden strings captured from the
reference itself and against the oracle.

Network built (greedy_strategy.py:70-295), in this order:
  L cores (as in qctn.adjacency_table) | L circuit states | Mx | R cores | R states
R cores are the same tensors (conjugated when complex) with the roles of the
in/out edge lists swapped and both lists reversed, but dims NOT permuted
(greedy_strategy.py:192-223, 764-822).  Then qubits are swept 0..n-1; at each
qubit every tensor touching it (plus circuit states hanging off those tensors)
is grouped by connectivity and each group becomes one einsum
(greedy_strategy.py:456-585, 615-664, 690-990).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from ..core.qctn import symbol_of


@dataclass(frozen=True)
class Operand:
    """Reference to one einsum operand.
    kind: 'core' (cores_dict[key]), 'core_conj' (conjugate if complex),
          'rcore' (right_cores_dict[key]), 'state' (circuit_states[key]),
          'mx' (measure_matrices[key]), 'tmp' (result of step `key`)."""
    kind: str
    key: object


@dataclass
class Step:
    qubit: int
    equation: str                 # remapped, exactly what the reference hands to torch.einsum
    operands: List[Operand]
    out: int                      # tmp id produced
    dims: Dict[str, int]          # symbol -> extent (batch symbols 'a', 'b' excluded)
    final: bool = False           # produced by the trailing "contract remaining" einsum
    # the same contraction with GLOBAL edge symbols (one symbol per network edge, before
    # the reference renames them per call): what the device plan is lowered from
    raw_subs: List[str] = field(default_factory=list)
    raw_out: str = ""
    raw_dims: Dict[str, int] = field(default_factory=dict)


@dataclass
class GreedySchedule:
    steps: List[Step]
    result: Operand
    batch: str                    # '', 'a' or 'ab' : batch symbols of the result
    equations: List[str] = field(default_factory=list)


class _Edge:
    __slots__ = ("nbr", "rank", "qubit", "sym", "xpart")

    def __init__(self, nbr, rank, qubit, xpart=False):
        self.nbr, self.rank, self.qubit, self.sym, self.xpart = nbr, rank, qubit, None, xpart

    def clone(self):
        e = _Edge(self.nbr, self.rank, self.qubit, self.xpart)
        e.sym = self.sym
        return e


class _Tensor:
    __slots__ = ("uid", "op", "ins", "outs", "side", "batch", "is_state")

    def __init__(self, uid, op, ins, outs, side, batch=""):
        self.uid, self.op, self.ins, self.outs, self.side, self.batch = uid, op, ins, outs, side, batch
        self.is_state = op.kind == "state"

    def touches(self, q):
        return any(e.qubit == q for e in self.ins) or any(e.qubit == q for e in self.outs)


def _edges_from(table_edges):
    return [_Edge(e["neighbor_idx"], e["edge_rank"], e["qubit_idx"], bool(e.get("is_cross_partition")))
            for e in table_edges]


def _fresh_symbols():
    i = 0
    while True:
        s = symbol_of(i)
        i += 1
        if s not in ("a", "b"):
            yield s


def _remap(eq: str) -> Tuple[str, Dict[str, str]]:
    """Rename symbols in order of first appearance to symbol_of(2), symbol_of(3), ...
    ('a' and 'b' are the batch letters and stay) -- greedy_strategy.py:886-904."""
    table = {"a": "a", "b": "b", ",": ",", "-": "-", ">": ">"}
    n = 2
    for ch in eq:
        if ch not in table:
            table[ch] = symbol_of(n)
            n += 1
    return "".join(table[c] for c in eq), table


def build_schedule(adjacency_table, nqubits: int, state_dims: Dict[int, int],
                   mx_info: Dict[int, Tuple[str, int, int]], right="symmetric",
                   right_table=None) -> GreedySchedule:
    """state_dims: qubit -> extent of the circuit state present on that qubit.
    mx_info: qubit -> (batch symbols '', 'a' or 'ab', rows, cols) for every non-None Mx.
    right: 'symmetric' | None | 'qctn' (then right_table is the other QCTN's table)."""
    net: List[_Tensor] = []
    left, lstate, mid, rgt, rstate = {}, {}, {}, {}, {}

    for t in adjacency_table:
        left[t["core_idx"]] = len(net)
        net.append(_Tensor(len(net), Operand("core", t["core_name"]),
                           _edges_from(t["in_edge_list"]), _edges_from(t["out_edge_list"]), "L"))
    for q in range(nqubits):
        if q in state_dims:
            lstate[q] = len(net)
            net.append(_Tensor(len(net), Operand("state", q), [], [_Edge(-1, state_dims[q], q)], "L"))
    for q in range(nqubits):
        if q in mx_info:
            bsym, rows, cols = mx_info[q]
            mid[q] = len(net)
            net.append(_Tensor(len(net), Operand("mx", q), [_Edge(-1, rows, q)], [_Edge(-1, cols, q)], "M", bsym))
    if isinstance(right, str) and right == "symmetric":
        for t in adjacency_table:
            rgt[t["core_idx"]] = len(net)
            net.append(_Tensor(len(net), Operand("core_conj", t["core_name"]),
                               _edges_from(t["out_edge_list"])[::-1], _edges_from(t["in_edge_list"])[::-1], "R"))
    elif isinstance(right, str) and right == "qctn":
        for t in right_table:
            rgt[t["core_idx"] + len(left)] = len(net)
            net.append(_Tensor(len(net), Operand("rcore", t["core_name"]),
                               _edges_from(t["in_edge_list"]), _edges_from(t["out_edge_list"]), "R"))
    elif right is not None:
        raise ValueError("Invalid right_qctn parameter.")
    for q in range(nqubits):
        if q in state_dims:
            rstate[q] = len(net)
            net.append(_Tensor(len(net), Operand("state", q), [_Edge(-1, state_dims[q], q)], [], "R"))

    # ---- wire open ends to states / measurements, renumber core neighbours ----
    def attach(t, edges, renumber, plugs, plug_side):
        for e in edges:
            if e.xpart:
                continue
            if e.nbr == -1:
                if e.qubit in plugs:
                    e.nbr = plugs[e.qubit]
                    getattr(net[e.nbr], plug_side)[0].nbr = t.uid
            elif e.nbr in renumber:
                e.nbr = renumber[e.nbr]

    for uid in left.values():
        attach(net[uid], net[uid].ins, left, lstate, "outs")
        attach(net[uid], net[uid].outs, left, mid, "ins")
    for uid in rgt.values():
        attach(net[uid], net[uid].ins, rgt, mid, "outs")
        attach(net[uid], net[uid].outs, rgt, rstate, "ins")

    # ---- one symbol per edge ----
    syms = _fresh_symbols()
    for t in net:
        for e in t.outs:
            if e.sym is not None:
                continue
            e.sym = next(syms)
            if e.nbr >= 0:
                for f in net[e.nbr].ins:
                    if f.nbr == t.uid and f.qubit == e.qubit:
                        f.sym = e.sym
                        break
    for t in net:
        for e in t.ins:
            if e.sym is None:
                e.sym = next(syms)

    # ---- sweep ----
    steps: List[Step] = []
    next_uid = len(net)
    for q in range(nqubits):
        here = [t for t in net if t.touches(q)]
        if not here:
            continue
        by_uid = {}
        for t in net:
            by_uid.setdefault(t.uid, t)
        hanging = []
        for t in here:
            for e in t.ins + t.outs:
                if e.nbr < 0:
                    continue
                o = by_uid.get(e.nbr)
                if o is not None and o.is_state and all(o is not x for x in here) and all(o is not x for x in hanging):
                    hanging.append(o)
        here = here + hanging
        fresh, dead, successor = [], set(), {}
        for gi, group in enumerate(_components(here)):
            if len(group) == 1 and not group[0].touches(q):
                continue
            merged, step = _group_step(group, q, len(steps))
            merged.uid = next_uid
            next_uid += 1
            steps.append(step)
            fresh.append(merged)
            for t in group:
                dead.add(t.uid)
                successor[t.uid] = merged.uid
        if not fresh:
            continue
        net = [t for t in net if t.uid not in dead] + fresh
        for t in net:
            for e in t.ins + t.outs:
                if e.nbr in successor:
                    e.nbr = successor[e.nbr]

    if not net:
        raise RuntimeError("No tensor left after contraction")
    if len(net) > 1:
        steps.append(_remaining_step(net, len(steps), steps))
        result, batch = Operand("tmp", steps[-1].out), steps[-1].equation.split("->")[1]
    else:
        result, batch = net[0].op, net[0].batch
    return GreedySchedule(steps, result, batch, [s.equation for s in steps])


def _components(members):
    """Connected components, ordered by first member (greedy_strategy.py:615-664)."""
    n = len(members)
    if n <= 1:
        return [list(members)] if n else []
    slot = {t.uid: i for i, t in enumerate(members)}
    root = list(range(n))

    def find(i):
        while root[i] != i:
            root[i] = root[root[i]]
            i = root[i]
        return i

    for i, t in enumerate(members):
        for e in t.outs + t.ins:
            j = slot.get(e.nbr) if e.nbr >= 0 else None
            if j is not None:
                ri, rj = find(i), find(j)
                if ri != rj:
                    root[ri] = rj
    comps: Dict[int, list] = {}
    for i in range(n):
        comps.setdefault(find(i), []).append(members[i])
    return list(comps.values())


def _group_step(group, q, step_id):
    inside = {t.uid for t in group}
    parts, ops, keep_in, keep_out, dims, bset = [], [], [], [], {}, set()

    def keep(e, bucket):
        internal = e.nbr >= 0 and e.nbr in inside
        if e.nbr == -1 or (not internal and e.qubit != q):
            bucket.append(e.clone())

    for t in group:
        ops.append(t.op)
        bset.update(t.batch)
        for e in t.ins + t.outs:
            dims[e.sym] = e.rank
        if t.side == "R":
            # dims of an R tensor follow the ORIGINAL core layout: the out list is the
            # reversed original in list, the in list the reversed original out list
            n_in, n_out = len(t.outs), len(t.ins)
            slots = [None] * (n_in + n_out)
            for i, e in enumerate(t.outs):
                slots[n_in - 1 - i] = e.sym
                keep(e, keep_out)
            for i, e in enumerate(t.ins):
                slots[n_in + n_out - 1 - i] = e.sym
                keep(e, keep_in)
            parts.append(t.batch + "".join(s for s in slots if s is not None))
        else:
            text = t.batch
            for e in t.ins:
                text += e.sym
                keep(e, keep_in)
            for e in t.outs:
                text += e.sym
                keep(e, keep_out)
            parts.append(text)
    batch = "".join(c for c in "ab" if c in bset)
    raw = ",".join(parts) + "->" + batch + "".join(e.sym for e in keep_in) + "".join(e.sym for e in keep_out)
    eq, table = _remap(raw)
    merged = _Tensor(-1, Operand("tmp", step_id), keep_in, keep_out, "M", batch)
    return merged, Step(q, eq, ops, step_id, {table[s]: d for s, d in dims.items()},
                        raw_subs=parts, raw_out=raw.split("->")[1], raw_dims=dict(dims))


def _remaining_step(net, step_id, steps):
    """Trailing einsum over whatever is left (greedy_strategy.py:993-1080).  The
    reference does NOT remap symbols here and keeps only batch symbols."""
    out_rank = {s.out: len(s.equation.split("->")[1]) for s in steps}
    parts, ops, dims, outsyms, true_subs, open_syms = [], [], {}, [], [], []
    for t in net:
        ops.append(t.op)
        for e in t.ins + t.outs:
            dims[e.sym] = e.rank
        # what the operand really looks like (the reference gets this wrong for merged
        # tensors, SURVEY defect D6); used to give disconnected networks their true value
        if t.op.kind == "tmp":
            true_subs.append(next(s for s in steps if s.out == t.op.key).raw_out)
        elif t.side == "R":
            n_in, n_out = len(t.outs), len(t.ins)
            slots = [None] * (n_in + n_out)
            for i, e in enumerate(t.outs):
                slots[n_in - 1 - i] = e.sym
            for i, e in enumerate(t.ins):
                slots[n_in + n_out - 1 - i] = e.sym
            true_subs.append(t.batch + "".join(slots))
        else:
            true_subs.append(t.batch + "".join(e.sym for e in t.ins) + "".join(e.sym for e in t.outs))
        if t.side == "M":
            if t.op.kind == "tmp":
                nb = out_rank[t.op.key] - 2
            else:
                nb = len(t.batch)
            text = ("a" if nb >= 1 else "") + ("b" if nb >= 2 else "")
            if t.ins:
                text += t.ins[0].sym
            if t.outs:
                text += t.outs[0].sym
        elif t.side == "R":
            n_in, n_out = len(t.outs), len(t.ins)
            slots = [None] * (n_in + n_out)
            for i, e in enumerate(t.outs):
                slots[n_in - 1 - i] = e.sym
            for i, e in enumerate(t.ins):
                slots[n_in + n_out - 1 - i] = e.sym
            text = "".join(s for s in slots if s is not None)
        else:
            text = "".join(e.sym for e in t.ins) + "".join(e.sym for e in t.outs)
        parts.append(text)
        for c in "ab":
            if c in text and c not in outsyms:
                outsyms.append(c)
    eq = ",".join(parts) + "->" + "".join(outsyms)
    count: Dict[str, int] = {}
    for sub in true_subs:
        for c in sub:
            count[c] = count.get(c, 0) + 1
    tb = "".join(c for c in "ab" if c in count)
    true_out = tb + "".join(c for sub in true_subs for c in sub if c not in "ab" and count[c] == 1)
    return Step(-1, eq, ops, step_id, dims, final=True, raw_subs=true_subs, raw_out=true_out, raw_dims=dict(dims))
