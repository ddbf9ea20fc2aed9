"""What happens on an UNFILTERED batch (no well-conditioned selection): the error budget of the
CUDA path against the float64 oracle, next to the reference's own float32 error on the same batch.

Samples whose amplitude nearly cancels have float32 values that are pure rounding noise in ANY
implementation.  The loss weights a sample's gradient contribution by 1/p, and is clamped at 1e-10
(engine_siamese.py:490-530): a sample whose TRUE value is below the noise floor contributes garbage
of size noise/p -- and whether its float32 value lands above or below the clamp is decided by
rounding.  So the full-batch gradient error is bounded in two parts:
  (a) values: for EVERY sample, |ours - float64| <= 10 x the reference's largest float32 value error
      on the batch (we are not noisier than the reference), and
  (b) gradients: on the samples whose float64 value is above 100 x that noise floor (where 1/p
      cannot amplify the noise beyond ~1e-2 relative per sample) our gradient error is <= 10 x the
      reference's own float32 gradient error on the same samples; the report printed by the test
      shows the full unfiltered numbers as well.
"""
import pytest
import torch

import tneq_b200
from oracle import qctn_oracle as oc
from helpers import make_case, clone_mx, upcast, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _graph(kind, n, K):
    H = tneq_b200.QCTNHelper
    g = H.generate_example_graph(n=n, graph_type="mps", dim_char=str(K))
    if kind == "merged":
        q = tneq_b200.QCTN(g)
        return tneq_b200.QCTN.merge(q, q).graph
    return g


def _gpu(graph, K, cores, states, mxs, sel=None):
    be = tneq_b200.BackendFactory.create_backend("b200", device=DEV, dtype="float32")
    eng = tneq_b200.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    q = tneq_b200.QCTN(graph, backend=be)
    for k, v in cores.items():
        q.cores_weights[k] = v.to(DEV).requires_grad_(True)
    st = [s.to(DEV) for s in states]

    def dev(ms):
        out = []
        for m in clone_mx(ms):
            t = m.tensor if isinstance(m, oc.TNT) else m
            t = t if sel is None else t[sel]
            out.append(tneq_b200.TNTensor(t.to(DEV), m.scale, m.log_scale) if isinstance(m, oc.TNT) else t.to(DEV))
        return out
    vals = eng.contract_with_compiled_strategy(q, st, dev(mxs))
    loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, st, dev(mxs))
    return vals.cpu(), loss.item(), [g.cpu() for g in grads]


def _sub(mxs, sel):
    out = []
    for m in clone_mx(mxs):
        if isinstance(m, oc.TNT):
            out.append(oc.TNT(m.tensor[sel].clone(), m.scale, m.log_scale))
        else:
            out.append(m[sel].clone())
    return out


@pytest.mark.parametrize("kind,n,K,B", [("mps", 16, 3, 4096), ("merged", 6, 3, 384)])
def test_unfiltered_batch_error_budget(kind, n, K, B, built_lib):
    graph = _graph(kind, n, K)
    names, table, nq, cores, states, mxs = make_case(graph, K, B, "float32", tnt=True, seed=21)
    c64 = {k: v.double() for k, v in cores.items()}
    s64 = [s.double() for s in states]
    m64 = lambda ms: [upcast(m, torch.float64) for m in clone_mx(ms)]
    truth = oc.forward(graph, c64, s64, m64(mxs))              # TNTensor results come back rescaled to scale 1
    ref32 = oc.forward(graph, cores, states, clone_mx(mxs))
    got, loss, grads = _gpu(graph, K, cores, states, mxs)
    # (a) values, every sample
    noise_ref = (ref32.double() - truth).abs().max().item()
    noise_ours = (got.double() - truth).abs().max().item()
    assert noise_ours <= 10 * noise_ref, (noise_ours, noise_ref)
    # full unfiltered gradient error (reported, not bounded: see the module docstring)
    tl, tg = oc.loss_and_grads(graph, c64, s64, m64(mxs))
    wl, wg = oc.loss_and_grads(graph, cores, states, clone_mx(mxs))
    full_ours = max(rel_err(g.double(), t) for g, t in zip(grads, tg))
    full_ref = max(rel_err(w.double(), t) for w, t in zip(wg, tg))
    # (b) gradients on the samples above 100 x the noise floor -- a data-independent criterion on the
    # float64 values, NOT a top-k selection
    sel = torch.nonzero(truth.abs() >= 100 * max(noise_ref, noise_ours)).flatten()
    frac = len(sel) / B
    print(f"\n[unfiltered {kind} n={n} B={B}] value noise: ours {noise_ours:.2e} reference {noise_ref:.2e} "
          f"(largest value {truth.abs().max():.2e}); full-batch gradient error vs float64: ours {full_ours:.2e} "
          f"reference {full_ref:.2e}; {len(sel)} samples above 100 x the noise floor")
    assert frac > 0.2, f"only {frac:.2f} of the batch is above the noise floor"
    tl_s, tg_s = oc.loss_and_grads(graph, c64, s64, m64(_sub(mxs, sel)))
    wl_s, wg_s = oc.loss_and_grads(graph, cores, states, _sub(mxs, sel))
    _, loss_s, grads_s = _gpu(graph, K, cores, states, mxs, sel=sel)
    err_ours = max(rel_err(g.double(), t) for g, t in zip(grads_s, tg_s))
    err_ref = max(rel_err(w.double(), t) for w, t in zip(wg_s, tg_s))
    print(f"[unfiltered {kind}] gradient error on the {len(sel)} samples above the noise floor: ours {err_ours:.2e} "
          f"reference {err_ref:.2e}")
    assert err_ours <= max(1e-5, 10 * err_ref), (err_ours, err_ref)
    assert abs(loss_s - tl_s.item()) <= max(1e-5 * abs(tl_s.item()), 10 * abs(wl_s.item() - tl_s.item()))
