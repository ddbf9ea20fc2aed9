"""numpy emulator of the device program format (TEST INFRASTRUCTURE ONLY).

Executes exactly the int64 blob that `libtneq_b200.so` receives
(contractor/vm_program.py documents the format), one op at a time, in float64
or float32.  It lets the CPU test-suite check the whole plan compiler
(symbolic greedy schedule -> pairwise graph -> reverse mode -> buffer
allocation -> op encoding) against the oracle without a GPU; the CUDA kernel
then only has to implement the op semantics faithfully, which the `-m gpu`
tests check against the oracle directly.  Never imported by the product.
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x544E5142323030
OP_WORDS = 24
SP_CONST, SP_FRAME, SP_GIN, SP_GOUT, SP_GACC = 0, 1, 2, 3, 4


class Parsed:
    pass


def parse(blob: np.ndarray) -> Parsed:
    b = np.asarray(blob, dtype=np.int64)
    assert b[0] == MAGIC and b[1] == 1
    p = Parsed()
    (p.dtype, p.n_in, p.n_out, p.const_elems, p.frame_elems, p.gacc_elems, p.n_prep, p.n_body, p.n_fin,
     p.n_itab, p.n_ftab, p.n_scalars, p.nb) = [int(v) for v in b[2:15]]
    at = 16
    slots = b[at: at + 2 * (p.n_in + p.n_out)].reshape(-1, 2)
    at += 2 * (p.n_in + p.n_out)
    p.in_slots, p.out_slots = slots[: p.n_in], slots[p.n_in:]
    nops = p.n_prep + p.n_body + p.n_fin
    ops = b[at: at + nops * OP_WORDS].reshape(nops, OP_WORDS)
    at += nops * OP_WORDS
    p.prep, p.body, p.fin = ops[: p.n_prep], ops[p.n_prep: p.n_prep + p.n_body], ops[p.n_prep + p.n_body:]
    p.itab = b[at: at + p.n_itab]
    at += p.n_itab
    p.ftab = b[at: at + p.n_ftab].view(np.float64)
    assert at + p.n_ftab == len(b)
    return p


def run(blob, inputs, nsamples, scalars=(0.0, 1.0), dtype=None, tile=5):
    """inputs: list (program order) of arrays; batched ones have shape [nsamples, elems].
    Returns the list of outputs (batched: [nsamples, elems]; shared: [elems])."""
    p = parse(blob)
    T = dtype or (np.float32 if p.dtype == 0 else np.float64)
    ins = [np.asarray(x, dtype=T).reshape(nsamples if p.in_slots[i][0] else 1, -1) for i, x in enumerate(inputs)]
    for i, x in enumerate(ins):
        assert x.shape[1] == p.in_slots[i][1], (i, x.shape, p.in_slots[i])
    outs = [np.zeros((nsamples if bt else 1, el), dtype=T) for bt, el in p.out_slots]
    const = np.zeros(p.const_elems, dtype=T)
    gacc = np.zeros(p.gacc_elems, dtype=T)

    def tab(off, n):
        return p.itab[off: off + n]

    def ftab(off, n):
        return p.ftab[off: off + n].astype(T)

    # mem(space, base, slot, rows) -> 2-D array view [rows_or_1, elems] + base
    def read(space, base, slot, offs, frame, rows):
        if space == SP_CONST:
            return const[base + offs][None, :]
        if space == SP_GACC:
            return gacc[base + offs][None, :]
        if space == SP_FRAME:
            return frame[:, base + offs]
        if space == SP_GIN:
            x = ins[slot]
            return x[rows][:, base + offs] if p.in_slots[slot][0] else x[:, base + offs]
        raise ValueError(space)

    def write(space, base, slot, offs, val, acc, frame, rows):
        if space == SP_CONST:
            tgt, r = const, None
        elif space == SP_GACC:
            tgt, r = gacc, None
        elif space == SP_FRAME:
            if acc:
                frame[:, base + offs] += val
            else:
                frame[:, base + offs] = val
            return
        elif space == SP_GOUT:
            o = outs[slot]
            if p.out_slots[slot][0]:
                if acc:
                    o[rows[:, None], (base + offs)[None, :]] += val
                else:
                    o[rows[:, None], (base + offs)[None, :]] = val
                return
            tgt, r = o[0], None
        else:
            raise ValueError(space)
        v = np.asarray(val).reshape(-1, len(offs))
        assert v.shape[0] == 1
        if acc:
            np.add.at(tgt, base + offs, v[0])
        else:
            tgt[base + offs] = v[0]

    def exec_op(w, frame, rows):
        code, acc = int(w[0]), int(w[1])
        if code == 1:      # LIN
            cnt, nt = int(w[6]), int(w[7])
            val = read(int(w[4]), int(w[5]), int(w[13]), tab(int(w[9]), cnt), frame, rows) * ftab(int(w[10]), cnt)
            if nt == 2:
                val = val + read(int(w[4]), int(w[5]), int(w[13]), tab(int(w[11]), cnt), frame, rows) * ftab(int(w[12]), cnt)
            write(int(w[2]), int(w[3]), int(w[14]), np.arange(cnt), val, acc, frame, rows)
        elif code == 2:    # GEMM
            nm, nn, nk = int(w[8]), int(w[9]), int(w[10])
            am, cm, ak = tab(int(w[11]), nm), tab(int(w[12]), nm), tab(int(w[13]), nk)
            bk, bn, cn = tab(int(w[14]), nk), tab(int(w[15]), nn), tab(int(w[16]), nn)
            A = read(int(w[4]), int(w[5]), int(w[17]), (am[:, None] + ak[None, :]).reshape(-1), frame, rows)
            B = read(int(w[6]), int(w[7]), int(w[18]), (bk[:, None] + bn[None, :]).reshape(-1), frame, rows)
            A = A.reshape(A.shape[0], nm, nk)
            B = B.reshape(B.shape[0], nk, nn)
            C = np.matmul(A, B).astype(T)
            C = C.reshape(C.shape[0], nm * nn)
            write(int(w[2]), int(w[3]), int(w[19]), (cm[:, None] + cn[None, :]).reshape(-1), C, acc, frame, rows)
        elif code == 3:    # RGEMM
            nm, nn, nk = int(w[8]), int(w[9]), int(w[10])
            am, dm, ak = tab(int(w[11]), nm), tab(int(w[12]), nm), tab(int(w[13]), nk)
            dn, gk, gn = tab(int(w[14]), nn), tab(int(w[15]), nk), tab(int(w[16]), nn)
            A = read(int(w[4]), int(w[5]), int(w[17]), (am[:, None] + ak[None, :]).reshape(-1), frame, rows)
            D = read(int(w[6]), int(w[7]), int(w[18]), (dm[:, None] + dn[None, :]).reshape(-1), frame, rows)
            A = A.reshape(A.shape[0], nm, nk)
            D = D.reshape(D.shape[0], nm, nn)
            G = np.einsum("smk,smn->kn", A, D).astype(T)
            np.add.at(gacc, int(w[3]) + (gk[:, None] + gn[None, :]).reshape(-1), G.reshape(-1))
        elif code == 4:    # SEED (fused loss): engine_siamese.py:490-530
            cplx, vb, dvb, lb = int(w[2]), int(w[3]), int(w[4]), int(w[5])
            log_scale, inv_count = T(scalars[int(w[6])]), T(scalars[int(w[7])])
            if cplx:
                vr, vi = frame[:, vb], frame[:, vb + 1]
                val = vr * vr + vi * vi
            else:
                val = frame[:, vb]
            clamped = np.maximum(val, T(1e-10))
            gacc[lb] += -np.sum(np.log(clamped) + log_scale) * inv_count
            dval = np.where(val >= T(1e-10), -inv_count / clamped, T(0)).astype(T)
            if cplx:
                frame[:, dvb] = dval * 2 * vr
                frame[:, dvb + 1] = dval * 2 * vi
            else:
                frame[:, dvb] = dval
        else:
            raise ValueError(code)

    for w in p.prep:
        exec_op(w, None, None)
    for t0 in range(0, nsamples, tile):
        rows = np.arange(t0, min(nsamples, t0 + tile))
        frame = np.full((len(rows), max(1, p.frame_elems)), np.nan, dtype=T)
        for w in p.body:
            exec_op(w, frame, rows)
    for w in p.fin:
        exec_op(w, None, None)
    return outs
