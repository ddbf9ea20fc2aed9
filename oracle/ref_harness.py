"""Import the REAL reference (`/root/reference/tneq_qc`) in the build container.

TEST INFRASTRUCTURE ONLY.  The reference is imported from /root/reference in the build
container and, on the GPU box (where that path does not exist), from the unmodified copy that
`__graft_entry__.build()` stages under the git-ignored baseline/_ref/ (`stage()` below).  Used by `oracle/make_golden.py` to (1) prove the
restatement in `qctn_oracle.py` bit-identical to the reference on CPU and
(2) generate the fixtures committed under tests/golden/.

Process-local shims -- /root/reference is never written:
  * `opt_einsum` stand-in exposing get_symbol, put on sys.path AFTER importing
    torch so torch.einsum keeps contracting left to right;
  * TNTensor.is_complex / TNTensor.conj (defect D2: the gradient path calls
    them, greedy_strategy.py:677-681, but tn_tensor.py does not define them).
"""

from __future__ import annotations

import contextlib
import io
import os
import sys

import torch  # noqa: F401  (must be imported before the shim is visible)

_HERE = os.path.dirname(os.path.abspath(__file__))
# where the reference package lives: the read-only checkout in the build container, else the
# git-ignored copy that build() stages under baseline/_ref/ so that it travels to the GPU box
STAGED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")
REF_ROOT = os.environ.get("TNEQ_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/tneq_qc") else STAGED)
_SHIM = os.path.join(_HERE, "_opt_einsum_shim")


def stage(src: str = "/root/reference") -> bool:
    """Copy the reference's Python package (unmodified) into baseline/_ref/ (git-ignored, not part of
    this repository's history; it only rides along to the GPU box).  Returns False where the
    reference checkout does not exist."""
    import shutil
    pkg = os.path.join(src, "tneq_qc")
    if not os.path.isdir(pkg):
        return False
    dst = os.path.join(STAGED, "tneq_qc")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(pkg, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "*.safetensors", "*.npz"))
    return True


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "tneq_qc"))


_loaded = None


def load():
    """Return a namespace with the reference classes used on the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    assert not torch.backends.opt_einsum.is_available()
    for p in (_SHIM, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):  # tneq_qc/config.py prints at import
        # import order matters: tneq_qc.core <-> tneq_qc.contractor are circular
        from tneq_qc.backends.backend_factory import BackendFactory
        from tneq_qc.core.engine_siamese import EngineSiamese
        from tneq_qc.contractor import StrategyCompiler, GreedyStrategy
        from tneq_qc.contractor.base import ContractionStrategy
        from tneq_qc.backends.backend_interface import ComputeBackend
        from tneq_qc.core.qctn import QCTN, QCTNHelper
        from tneq_qc.core.tn_tensor import TNTensor
        from tneq_qc.optim.optimizer import Optimizer
    TNTensor.is_complex = lambda s: s.tensor.is_complex()
    TNTensor.conj = lambda s: TNTensor(s.tensor.conj(), s.scale, s.log_scale)

    class NS:
        pass

    ns = NS()
    ns.BackendFactory, ns.EngineSiamese, ns.QCTN, ns.QCTNHelper = BackendFactory, EngineSiamese, QCTN, QCTNHelper
    ns.TNTensor, ns.Optimizer, ns.StrategyCompiler, ns.GreedyStrategy = TNTensor, Optimizer, StrategyCompiler, GreedyStrategy
    ns.ContractionStrategy, ns.ComputeBackend = ContractionStrategy, ComputeBackend
    _loaded = ns
    return ns


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


@contextlib.contextmanager
def spy_einsum(log):
    """Record (equation, operand shapes) of every torch.einsum call."""
    orig = torch.einsum

    def wrapped(eq, *ops):
        log.append((eq, [tuple(o.shape) for o in ops]))
        return orig(eq, *ops)

    torch.einsum = wrapped
    try:
        yield
    finally:
        torch.einsum = orig


def make_engine(dtype="float32", K=3):
    ns = load()
    with quiet():
        be = ns.BackendFactory.create_backend("pytorch", device="cpu", dtype=dtype)
        eng = ns.EngineSiamese(backend=be, strategy_mode="balanced", mx_K=K)
    return be, eng


def ref_forward(graph, cores, states, mxs, dtype="float32", log=None):
    """Reference forward on CPU.  `cores`: name -> tensor (set via cores_weights)."""
    ns = load()
    be, eng = make_engine(dtype)
    with quiet():
        q = ns.QCTN(graph, backend=be)
        for k, v in cores.items():
            q.cores_weights[k] = v
        with spy_einsum(log if log is not None else []):
            out = eng.contract_with_compiled_strategy(q, states, mxs)
    return out


def ref_loss_and_grads(graph, cores, states, mxs, dtype="float32", log=None):
    ns = load()
    be, eng = make_engine(dtype)
    with quiet():
        q = ns.QCTN(graph, backend=be)
        for k, v in cores.items():
            t = v.detach().clone().requires_grad_(True)
            q.cores_weights[k] = t
        with spy_einsum(log if log is not None else []):
            loss, grads = eng.contract_with_compiled_strategy_for_gradient(q, states, mxs)
    return loss, list(grads)
