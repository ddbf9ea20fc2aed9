"""Device program ("VM program") lowered from a contraction graph.

A program is what `libtneq_b200.so` executes (include/tneq_b200.h, tnq_plan_*):
three op lists over five memory spaces, plus integer offset tables and float
coefficient tables.  It is encoded as one flat int64 array (`to_blob`).

Sections
  PREP  batch-independent ops, run once per call by one CTA: cores contracted
        with circuit states, conjugation / 2x2-real expansion of complex cores.
        Results live in the CONST pool.
  BODY  per-sample ops.  The batch is cut into tiles of S samples; a CTA keeps
        a tile's working set (the FRAME) in shared memory, structure-of-arrays
        [element][sample], walks the whole op list for that tile (the complete
        qubit sweep and, for training, the complete reverse sweep), and only
        then takes the next tile.  Gradients of shared tensors are summed over
        samples into a per-CTA accumulator (GACC) by warp-shuffle reductions.
  FIN   after a deterministic cross-CTA reduction of GACC: the reverse of PREP
        (gradients w.r.t. the caller's core tensors) and the loss.

Spaces: 0 CONST (pool of prepared shared tensors)   1 FRAME (per-sample scratch)
        2 GIN  (caller input slot)                   3 GOUT  (caller output slot)
        4 GACC (accumulators of shared adjoints + loss)

Ops (24 int64 words each; T = tables are offsets into itab / ftab):
  LIN   dst[j] (=|+=) c0[j]*src[s0[j]] + c1[j]*src[s1[j]]           j < count
  GEMM  C[cm[r]+cn[c]] (=|+=) sum_k A[am[r]+ak[k]] * B[bk[k]+bn[c]]   r < nm, c < nn
        (w[20] = TN column tile; w[22] = 1: B is a packed PREP copy [k][c/TN][TN padded to 16 B],
         row length w[21], so bk[k] = k*ldb and a thread's TN columns are contiguous and aligned)
  RGEMM G[gk[i]+gn[c]] += sum_{samples in tile} sum_r A[am[r]+ak[i]] * D[dm[r]+dn[c]]
  SEED  fused loss: value -> d(loss)/d(value), loss partial  (engine_siamese.py:490-530)
All tensors are compact row-major in their index order, so offsets are plain
dot products of multi-indices with strides; they are precomputed here so the
kernel does no index arithmetic beyond table lookups.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .cgraph import CGraph, Node

SP_CONST, SP_FRAME, SP_GIN, SP_GOUT, SP_GACC = 0, 1, 2, 3, 4
OP_LIN, OP_GEMM, OP_RGEMM, OP_SEED = 1, 2, 3, 4
OP_WORDS = 24
MAGIC = 0x544E5142323030  # "TNQB200"
VERSION = 1
SC_LOG_SCALE, SC_INV_COUNT = 0, 1   # runtime scalars


@dataclass
class Slot:
    key: Tuple[str, object]
    batched: bool
    elems: int              # real elements (per sample if batched)


@dataclass
class VMProgram:
    dtype: str                               # 'f32' | 'f64'
    inputs: List[Slot] = field(default_factory=list)
    outputs: List[Slot] = field(default_factory=list)
    const_elems: int = 0
    frame_elems: int = 0
    gacc_elems: int = 0
    prep: List[List[int]] = field(default_factory=list)
    body: List[List[int]] = field(default_factory=list)
    fin: List[List[int]] = field(default_factory=list)
    itab: List[int] = field(default_factory=list)
    ftab: List[float] = field(default_factory=list)
    nb: int = 1                              # samples per 'a' entry (2 for (B,2,K,K) measurements)
    flops_per_sample: float = 0.0
    meta: Dict[str, object] = field(default_factory=dict)

    def input_index(self, key):
        for i, s in enumerate(self.inputs):
            if s.key == key:
                return i
        raise KeyError(key)

    def output_index(self, key):
        for i, s in enumerate(self.outputs):
            if s.key == key:
                return i
        raise KeyError(key)

    def to_blob(self) -> np.ndarray:
        head = [MAGIC, VERSION, 0 if self.dtype == "f32" else 1, len(self.inputs), len(self.outputs),
                self.const_elems, self.frame_elems, self.gacc_elems, len(self.prep), len(self.body),
                len(self.fin), len(self.itab), len(self.ftab), 2, self.nb, 0]
        words = list(head)
        for s in self.inputs + self.outputs:
            words += [1 if s.batched else 0, s.elems]
        for op in self.prep + self.body + self.fin:
            assert len(op) == OP_WORDS
            words += op
        blob = np.array(words, dtype=np.int64)
        it = np.array(self.itab, dtype=np.int64)
        ft = np.array(self.ftab, dtype=np.float64).view(np.int64)
        return np.concatenate([blob, it, ft])


class _Tables:
    def __init__(self, prog: VMProgram):
        self.prog = prog
        self.icache: Dict[bytes, int] = {}
        self.fcache: Dict[bytes, int] = {}

    def ints(self, arr) -> int:
        a = np.ascontiguousarray(arr, dtype=np.int64)
        k = a.tobytes()
        if k not in self.icache:
            self.icache[k] = len(self.prog.itab)
            self.prog.itab.extend(int(v) for v in a)
        return self.icache[k]

    def floats(self, arr) -> int:
        a = np.ascontiguousarray(arr, dtype=np.float64)
        k = a.tobytes()
        if k not in self.fcache:
            self.fcache[k] = len(self.prog.ftab)
            self.prog.ftab.extend(float(v) for v in a)
        return self.fcache[k]


def _strides(g: CGraph, idx):
    st, s = {}, 1
    for i in reversed(idx):
        st[i] = s
        s *= g.dims[i]
    return st


def _offsets(g: CGraph, order, strides):
    """Offsets of all multi-indices over `order` (row-major enumeration) for a
    tensor with the given strides."""
    if not order:
        return np.zeros(1, dtype=np.int64)
    shp = [g.dims[i] for i in order]
    grid = np.indices(shp).reshape(len(shp), -1)
    off = np.zeros(grid.shape[1], dtype=np.int64)
    for k, i in enumerate(order):
        off += grid[k] * strides[i]
    return off


def _pick_tile(extent: int, choices) -> int:
    """Largest register-tile edge from `choices` that divides `extent`."""
    for c in choices:
        if extent % c == 0:
            return c
    return 1


def lower(g: CGraph, mode: str, dtype: str, nb: int = 1, emit_values: bool = True) -> VMProgram:
    """mode: 'fwd'   result only
             'bwd'   forward recomputation + reverse sweep seeded by the caller's grad (autograd route)
             'train' forward + fused loss + reverse sweep (returns loss, grads[, values])"""
    prog = VMProgram(dtype=dtype, nb=nb)
    tabs = _Tables(prog)
    nodes = g.nodes

    # ---- caller slots -------------------------------------------------------
    in_slot: Dict[Tuple[str, object], int] = {}
    for key, ids in g.inputs.items():
        n = nodes[ids[0]]
        in_slot[key] = len(prog.inputs)
        prog.inputs.append(Slot(key, n.batched, g.size(n.idx)))
    res = nodes[g.result]
    out_slot: Dict[Tuple[str, object], int] = {}

    def add_out(key, batched, elems):
        out_slot[key] = len(prog.outputs)
        prog.outputs.append(Slot(key, batched, elems))
        return out_slot[key]

    # ---- where every node lives ---------------------------------------------
    space: Dict[int, int] = {}
    base: Dict[int, int] = {}
    slot_of: Dict[int, int] = {}
    for n in nodes:
        if n.kind == "input":
            space[n.id], base[n.id] = SP_GIN, 0
            slot_of[n.id] = in_slot[(n.operand.kind, n.operand.key)]

    def consumers_of(nid):
        return [m for m in nodes if m.kind in ("contract", "lin", "seed") and (m.p == nid or m.q == nid)]

    # shared forward nodes -> CONST pool ; accumulators -> GACC
    for n in nodes:
        if n.kind == "input":
            continue
        if n.is_accum:
            space[n.id], base[n.id] = SP_GACC, prog.gacc_elems
            prog.gacc_elems += g.size(n.idx)
        elif not n.batched and n.acc_into < 0:
            space[n.id], base[n.id] = SP_CONST, prog.const_elems
            prog.const_elems += g.size(n.idx)
    loss_base = -1
    if mode == "train":
        loss_base = prog.gacc_elems
        prog.gacc_elems += 1

    # staging copies of caller inputs that feed a contraction
    staged: Dict[int, Tuple[int, int]] = {}     # input node id -> (space, base) ; FRAME bases are patched later

    # ---- BODY op list with symbolic FRAME buffers ---------------------------
    # each entry: dict(op=..., reads=[buf ids], writes=buf id or None)
    body_ops: List[dict] = []
    frame_size: Dict[object, int] = {}          # buffer id -> elems

    def frame_buf(bid, elems):
        frame_size[bid] = elems
        return bid

    def operand_loc(n: Node, for_body: bool):
        """(space, base-or-bufid, slot) of a tensor as a contraction operand."""
        if n.kind == "input":
            if n.batched:
                bid = ("stage", n.id)
                if bid not in frame_size:
                    frame_buf(bid, g.size(n.idx))
                    cnt = g.size(n.idx)
                    body_ops.append(dict(kind="lin", acc=0, dst=(SP_FRAME, bid, -1), src=(SP_GIN, 0, slot_of[n.id]),
                                         count=cnt, sf=np.arange(cnt)[:, None], cf=np.ones((cnt, 1))))
                return (SP_FRAME, bid, -1)
            if for_body:
                if n.id not in staged:
                    cnt = g.size(n.idx)
                    staged[n.id] = (SP_CONST, prog.const_elems)
                    prog.const_elems += cnt
                    prep_ops.append(dict(kind="lin", acc=0, dst=(SP_CONST, staged[n.id][1], -1),
                                         src=(SP_GIN, 0, slot_of[n.id]), count=cnt,
                                         sf=np.arange(cnt)[:, None], cf=np.ones((cnt, 1))))
                return (SP_CONST, staged[n.id][1], -1)
            return (SP_GIN, 0, slot_of[n.id])
        if space.get(n.id) in (SP_CONST, SP_GACC):
            return (space[n.id], base[n.id], -1)
        return (SP_FRAME, ("node", n.id), -1)

    prep_ops: List[dict] = []
    fin_ops: List[dict] = []
    packed_b: Dict[object, int] = {}

    def gemm_tables(c_idx, a: Node, b: Node):
        sa, sb, sc = _strides(g, a.idx), _strides(g, b.idx), _strides(g, c_idx)
        k_idx = [i for i in a.idx if i in b.idx]
        m_idx = [i for i in c_idx if i in a.idx and i not in b.idx]
        n_idx = [i for i in c_idx if i in b.idx and i not in a.idx]
        assert sorted(m_idx + n_idx) == sorted(c_idx), "output index that belongs to neither operand"
        assert all((i in c_idx) != (i in k_idx) for i in a.idx) and all((i in c_idx) != (i in k_idx) for i in b.idx)
        return dict(nm=g.size(m_idx), nn=g.size(n_idx), nk=g.size(k_idx),
                    am=_offsets(g, m_idx, sa), cm=_offsets(g, m_idx, sc), ak=_offsets(g, k_idx, sa),
                    bk=_offsets(g, k_idx, sb), bn=_offsets(g, n_idx, sb), cn=_offsets(g, n_idx, sc))

    def dst_loc(n: Node):
        if n.acc_into >= 0:
            t = nodes[n.acc_into]
            return (space[t.id], base[t.id], -1), 1
        if space.get(n.id) == SP_CONST:
            return (SP_CONST, base[n.id], -1), 0
        frame_buf(("node", n.id), g.size(n.idx))
        return (SP_FRAME, ("node", n.id), -1), 0

    for n in nodes:
        if n.kind == "input" or n.is_accum:
            continue
        if mode == "fwd" and n.role == "adj":
            continue
        p = nodes[n.p] if n.p >= 0 else None
        q = nodes[n.q] if n.q >= 0 else None
        if n.kind == "seed":
            frame_buf(("node", n.id), g.size(n.idx))
            body_ops.append(dict(kind="seed", cplx=1 if n.cplx else 0, v=(SP_FRAME, ("node", p.id), -1),
                                 dv=(SP_FRAME, ("node", n.id), -1)))
            continue
        if n.kind == "lin":
            in_body = n.batched or p.batched
            if n.batched:
                src = operand_loc(p, True) if p.kind != "input" else (SP_GIN, 0, slot_of[p.id])
            else:
                src = (space[p.id], base[p.id], slot_of.get(p.id, -1)) if p.kind != "input" else (SP_GIN, 0, slot_of[p.id])
            dst, acc = dst_loc(n)
            sf_, cf_ = g.lin_tables(n)
            op = dict(kind="lin", acc=acc, dst=dst, src=src, count=g.size(n.idx), sf=sf_, cf=cf_)
            if n.batched:
                body_ops.append(op)
            elif n.role == "adj":
                fin_ops.append(op)
            else:
                prep_ops.append(op)
            continue
        # contract
        if n.reduce_batch:
            # G[me.idx] += sum_samples sum_shared  P * Q   (both batched)
            a_loc, d_loc = operand_loc(p, True), operand_loc(q, True)
            sa, sd, sg = _strides(g, p.idx), _strides(g, q.idx), _strides(g, n.idx)
            m_idx = [i for i in p.idx if i in q.idx]
            k_idx = [i for i in n.idx if i in p.idx]
            c_idx = [i for i in n.idx if i in q.idx]
            assert not (set(k_idx) & set(c_idx)) and sorted(k_idx + c_idx) == sorted(n.idx)
            assert n.acc_into >= 0
            t = nodes[n.acc_into]
            body_ops.append(dict(kind="rgemm", g=(SP_GACC, base[t.id], -1), a=a_loc, d=d_loc,
                                 rk=_pick_tile(g.size(k_idx), (4, 3, 2, 1)), rn=_pick_tile(g.size(c_idx), (4, 3, 2, 1)),
                                 nm=g.size(m_idx), nk=g.size(k_idx), nn=g.size(c_idx),
                                 am=_offsets(g, m_idx, sa), dm=_offsets(g, m_idx, sd),
                                 ak=_offsets(g, k_idx, sa), dn=_offsets(g, c_idx, sd),
                                 gk=_offsets(g, k_idx, sg), gn=_offsets(g, c_idx, sg)))
            continue
        if n.batched:
            # A = the batched operand (the bigger one if both are), B = the other
            if p.batched and q.batched:
                a, b = (p, q) if g.size(p.idx) >= g.size(q.idx) else (q, p)
            else:
                a, b = (p, q) if p.batched else (q, p)
            tb = gemm_tables(n.idx, a, b)
            tn = _pick_tile(tb["nn"], (9, 8, 5, 4, 3, 2, 1))
            a_loc = operand_loc(a, True)
            extra = dict(tn=tn, ldb=0, packed=0)
            if b.batched:
                b_loc = operand_loc(b, True)
            else:
                # shared B: a PREP copy packed as [k][column group][TN padded to 16 bytes] so that the
                # kernel fetches a thread's TN columns with aligned vector loads (broadcast)
                vec = 4 if dtype == "f32" else 2
                tnp = -(-tn // vec) * vec
                ldb = (tb["nn"] // tn) * tnp
                pkey = (b.id, tuple(tb["bk"]), tuple(tb["bn"]), tn)
                if pkey not in packed_b:
                    prog.const_elems = -(-prog.const_elems // 4) * 4
                    pbase = prog.const_elems
                    prog.const_elems += tb["nk"] * ldb
                    sf = np.zeros(tb["nk"] * ldb, dtype=np.int64)
                    cf = np.zeros(tb["nk"] * ldb, dtype=np.float64)
                    for k in range(tb["nk"]):
                        for c in range(tb["nn"]):
                            at = k * ldb + (c // tn) * tnp + c % tn
                            sf[at], cf[at] = tb["bk"][k] + tb["bn"][c], 1.0
                    prep_ops.append(dict(kind="lin", acc=0, dst=(SP_CONST, pbase, -1), src=operand_loc(b, False),
                                         count=tb["nk"] * ldb, sf=sf[:, None], cf=cf[:, None]))
                    packed_b[pkey] = pbase
                b_loc = (SP_CONST, packed_b[pkey], -1)
                tb["bk"] = np.arange(tb["nk"], dtype=np.int64) * ldb
                tb["bn"] = np.array([(c // tn) * tnp + c % tn for c in range(tb["nn"])], dtype=np.int64)
                extra = dict(tn=tn, ldb=ldb, packed=1)
            dst, acc = dst_loc(n)
            body_ops.append(dict(kind="gemm", acc=acc, c=dst, a=a_loc, b=b_loc, **tb, **extra))
        else:
            a_loc, b_loc = operand_loc(p, False), operand_loc(q, False)
            dst, acc = dst_loc(n)
            op = dict(kind="gemm", acc=acc, c=dst, a=a_loc, b=b_loc, **gemm_tables(n.idx, p, q))
            (fin_ops if n.role == "adj" else prep_ops).append(op)

    # ---- results ---------------------------------------------------------------
    res_elems = g.size(res.idx)
    res_loc = (SP_FRAME, ("node", res.id), -1)
    if mode == "fwd" or (mode == "train" and emit_values):
        so = add_out(("result", 0), True, res_elems)
        # place the copy right after the op that produces the result
        pos = next(i for i, op in enumerate(body_ops)
                   if op.get("c", op.get("dst", (None, None)))[:2] == (SP_FRAME, ("node", res.id))) + 1
        body_ops.insert(pos, dict(kind="lin", acc=0, dst=(SP_GOUT, 0, so), src=res_loc, count=res_elems,
                                  sf=np.arange(res_elems)[:, None], cf=np.ones((res_elems, 1))))
    if mode in ("bwd", "train"):
        for key, nid in g.grads.items():
            cnt = g.size(nodes[nid].idx)
            so = add_out(("grad",) + key, False, cnt)
            fin_ops.append(dict(kind="lin", acc=0, dst=(SP_GOUT, 0, so), src=(SP_GACC, base[nid], -1), count=cnt,
                                sf=np.arange(cnt)[:, None], cf=np.ones((cnt, 1))))
    if mode == "train":
        so = add_out(("loss", 0), False, 1)
        fin_ops.append(dict(kind="lin", acc=0, dst=(SP_GOUT, 0, so), src=(SP_GACC, loss_base, -1), count=1,
                            sf=np.zeros((1, 1), dtype=np.int64), cf=np.ones((1, 1))))
        for op in body_ops:
            if op["kind"] == "seed":
                op["loss"] = loss_base
                op["vout"] = -1

    # ---- FRAME allocation by liveness (first fit) -----------------------------------
    def bufs(loc):
        return [loc[1]] if loc[0] == SP_FRAME else []

    first, last = {}, {}
    for i, op in enumerate(body_ops):
        touched = []
        for k in ("dst", "src", "c", "a", "b", "d", "v", "dv"):
            if k in op and isinstance(op[k], tuple):
                touched += bufs(op[k])
        for bfr in touched:
            first.setdefault(bfr, i)
            last[bfr] = i
    placed: Dict[object, int] = {}
    active: List[Tuple[int, int, object]] = []   # (base, size, buf)
    order = sorted(first, key=lambda b_: first[b_])
    peak = 0
    for bfr in order:
        t0 = first[bfr]
        active = [x for x in active if last[x[2]] >= t0]
        size = frame_size[bfr]
        active.sort()
        at = 0
        for b0, sz, _ in active:
            if at + size <= b0:
                break
            at = max(at, b0 + sz)
        placed[bfr] = at
        active.append((at, size, bfr))
        peak = max(peak, at + size)
    prog.frame_elems = peak

    # ---- encode ------------------------------------------------------------------------
    def loc3(loc):
        sp, b, sl = loc
        if sp == SP_FRAME:
            b = placed[b]
        return sp, int(b), int(sl)

    def encode(op) -> List[int]:
        w = [0] * OP_WORDS
        if op["kind"] == "lin":
            s, c = np.asarray(op["sf"]), np.asarray(op["cf"], dtype=np.float64)
            nt = s.shape[1]
            w[0], w[1] = OP_LIN, int(op["acc"])
            w[2], w[3], w[14] = loc3(op["dst"])
            w[4], w[5], w[13] = loc3(op["src"])
            w[6], w[7] = int(op["count"]), nt
            w[9], w[10] = tabs.ints(s[:, 0]), tabs.floats(c[:, 0])
            if nt == 2:
                w[11], w[12] = tabs.ints(s[:, 1]), tabs.floats(c[:, 1])
        elif op["kind"] == "gemm":
            w[0], w[1] = OP_GEMM, int(op["acc"])
            w[2], w[3], w[19] = loc3(op["c"])
            w[4], w[5], w[17] = loc3(op["a"])
            w[6], w[7], w[18] = loc3(op["b"])
            w[8], w[9], w[10] = int(op["nm"]), int(op["nn"]), int(op["nk"])
            w[11], w[12], w[13] = tabs.ints(op["am"]), tabs.ints(op["cm"]), tabs.ints(op["ak"])
            w[14], w[15], w[16] = tabs.ints(op["bk"]), tabs.ints(op["bn"]), tabs.ints(op["cn"])
            w[20], w[21], w[22] = int(op.get("tn", 0)), int(op.get("ldb", 0)), int(op.get("packed", 0))
        elif op["kind"] == "rgemm":
            w[0], w[1] = OP_RGEMM, 1
            w[2], w[3], _ = loc3(op["g"])
            w[4], w[5], w[17] = loc3(op["a"])
            w[6], w[7], w[18] = loc3(op["d"])
            w[8], w[9], w[10] = int(op["nm"]), int(op["nn"]), int(op["nk"])
            w[11], w[12], w[13] = tabs.ints(op["am"]), tabs.ints(op["dm"]), tabs.ints(op["ak"])
            w[14], w[15], w[16] = tabs.ints(op["dn"]), tabs.ints(op["gk"]), tabs.ints(op["gn"])
            w[20], w[21] = int(op["rk"]), int(op["rn"])
        elif op["kind"] == "seed":
            w[0] = OP_SEED
            w[2] = int(op["cplx"])
            w[3] = loc3(op["v"])[1]
            w[4] = loc3(op["dv"])[1]
            w[5] = int(op["loss"])
            w[6], w[7] = SC_LOG_SCALE, SC_INV_COUNT
            w[8] = int(op.get("vout", -1))
        return w

    prog.prep = [encode(op) for op in prep_ops]
    prog.body = [encode(op) for op in body_ops]
    prog.fin = [encode(op) for op in fin_ops]
    prog.flops_per_sample = sum(2.0 * op["nm"] * op["nn"] * op["nk"] for op in body_ops if op["kind"] in ("gemm", "rgemm"))
    prog.meta = dict(n_prep=len(prog.prep), n_body=len(prog.body), n_fin=len(prog.fin),
                     result_elems=res_elems, result_cplx=bool(res.cplx))
    return prog
