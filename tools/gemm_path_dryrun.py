"""Dry run of the large-bond route (contractor/gemm_path.py) on META tensors: no GPU, no arithmetic.
Prints every permute (bytes moved) and GEMM (M, N, K, batch) of one forward + backward pass of an n-qubit MPS
network of bond dimension chi (complex64), so that the data movement of a plan can be read without a device.
    python tools/gemm_path_dryrun.py [n] [chi] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tneq_b200
from tneq_b200.contractor import gemm_path
from tneq_b200.contractor.plan import ContractionPlan, signature_of


class FakeLib:
    """expand / fold are called on the library directly: log the elements they write"""

    def __init__(self, log):
        self.log = log

    def _elementwise(self, name, dims, k):
        n = 1
        for i in range(k):
            n *= dims[i]
        self.log.append((name, n * 4, [dims[i] for i in range(k)], [], 0))
        return 0

    def tnq_cplx_expand_f32(self, src, dst, k, dims, *rest):
        return self._elementwise("expand", dims, k)

    def tnq_cplx_fold_f32(self, src, dst, k, dims, *rest):
        return self._elementwise("fold", dims, k)

    def tnq_fold_vec_f32(self, P, Q, out, A, D, C, stream):
        self.log.append(("foldvec", (A * D * C * 2 + A * C * 2) * 4 / 2, [A, D, C], [], 0))
        return 0

    def tnq_outer_acc_f32(self, P, Q, T, A, D, C, stream):
        self.log.append(("outeracc", (A * D * C * 2 * 2 + A * C * 2) * 4 / 2, [A, D, C], [], 0))
        return 0

    def tnq_gemm_tf32x3_bk(self, A, a_mn, a_tiles, B, b_mn, b_tiles, C, M, N, batch, Kin, stream):
        self.log.append(("gemm", 2.0 * M * N * batch * Kin, M, N, batch * Kin, 1, ("bk", a_mn, b_mn)))
        print(f"  batch-into-K GEMM in place: M {M} N {N} K {batch * Kin} a_mn {a_mn} b_mn {b_mn}")
        return 0

    def tnq_gemm_tf32x3_view(self, A, R1, R0, sR1, sR0, K1, K0, sK1, B, ldb, C, ldc, N, stream):
        ok = K0 % 32 == 0 and (R0 % 128 == 0 or 128 % R0 == 0) and K1 * K0 > 256 and not (sR1 % 4 or sR0 % 4 or sK1 % 4)
        if ok:
            self.log.append(("gemm", 2.0 * R1 * R0 * N * K1 * K0, R1 * R0, N, K1 * K0, 1, ("view", R1, R0, sR1, sR0, K1, K0, sK1)))
        return 0 if ok else -2


class DryRunner(gemm_path.GemmPathRunner):
    def __init__(self, graph):
        self.g, self.device, self.flops = graph, torch.device("meta"), 0.0
        self.log = []
        self.lib = FakeLib(self.log)

    def _stream(self):
        return None

    def _permute(self, src, src_strides, out_dims, vec, conj):
        n = 1
        for d in out_dims:
            n *= d
        self.log.append(("permute", n * 8, list(out_dims), list(src_strides), vec))
        return torch.empty(tuple(out_dims), dtype=torch.float32, device="meta")

    def _gemm_view(self, A, view, Bm, C, N, ldb, ldc):
        rc = self.lib.tnq_gemm_tf32x3_view(None, *view, None, ldb, None, ldc, N, None)
        if rc == 0:
            self.flops += 2.0 * view[0] * view[1] * N * view[4] * view[5]
        return rc == 0

    def _gemm(self, A, B, C, M, N, K, lda, ldb, ldc, batch=1, sA=0, sB=0, sC=0, accumulate=False):
        self.log.append(("gemm", 2.0 * M * N * K * batch, M, N, K, batch, getattr(self, "_cur", None)))
        self.flops += 2.0 * M * N * K * batch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    graph = tneq_b200.QCTNHelper.generate_example_graph(n=n, graph_type="mps", dim_char=str(chi))
    q = tneq_b200.QCTN(graph)
    states = [torch.empty(chi, dtype=torch.complex64, device="meta") for _ in range(n)]
    mxs = [torch.empty(B, chi, chi, dtype=torch.complex64, device="meta") for _ in range(n)]
    sd, mi = signature_of(q.nqubits, states, mxs)
    plan = ContractionPlan(q.adjacency_table, q.nqubits, {c: q.core_shape(c) for c in q.cores}, sd, mi, "complex64")
    g = plan.graph("bwd")
    r = DryRunner(g)
    inputs = {}
    for key, ids in g.inputs.items():
        node = g.nodes[ids[0]]
        size = g.size(node.idx)
        inputs[key] = torch.empty((B, size) if node.batched else (size,), dtype=torch.float32, device="meta")
    seed = torch.empty(B * g.size(g.nodes[g.result].idx), dtype=torch.float32, device="meta")

    orig_lin, orig_con = r._lin, r._contract

    def lin(nn, val, lay, NS):
        r._cur = ("lin", nn.lin_kind, nn.id, nn.role)
        k = len(r.log)
        out = orig_lin(nn, val, lay, NS)
        for e in r.log[k:]:
            print(f"  node {nn.id:4d} {nn.role:4s} lin/{nn.lin_kind:8s} {e[0]:8s} {e[1] / 1e6:10.1f} MB  src layout {lay[g.nodes[nn.p].id]} -> {out[1]}")
        return out

    def con(nn, val, lay, NS):
        k = len(r.log)
        p, qn = g.nodes[nn.p], g.nodes[nn.q]
        out = orig_con(nn, val, lay, NS)
        for e in r.log[k:]:
            if e[0] != "gemm":
                print(f"  node {nn.id:4d} {nn.role:4s} contract     {e[0]:8s} {e[1] / 1e6:10.1f} MB  dims {e[2]} strides {e[3]} vec {e[4]}")
            else:
                print(f"  node {nn.id:4d} {nn.role:4s} contract     GEMM     {e[1] / 1e9:10.2f} GF  M {e[2]} N {e[3]} K {e[4]} batch {e[5]}"
                      f"  p{'(b)' if p.batched else ''} {lay[p.id]} x q{'(b)' if qn.batched else ''} {lay[qn.id]} -> {out[1]} reduce_batch={nn.reduce_batch}{' VIEW ' + str(e[6][1:]) if isinstance(e[6], tuple) and e[6] and e[6][0] == 'view' else ''}")
        return out

    r._lin, r._contract = lin, con
    r.run(inputs, B, 1, with_adjoint=True, seed=seed)
    pb = sum(e[1] for e in r.log if e[0] != "gemm")
    print(f"total: {sum(1 for e in r.log if e[0] != 'gemm')} permutes / expansions / folds moving {2 * pb / 1e9:.2f} GB (read+write), "
          f"{sum(1 for e in r.log if e[0] == 'gemm')} GEMMs, {r.flops / 1e12:.3f} TFLOP")


if __name__ == "__main__":
    main()
