"""Plug tneq_b200 into an importable copy of the REFERENCE package (tneq_qc).

    import tneq_qc, tneq_b200.reference_plugin as plug
    plug.register(tneq_qc)
    backend = tneq_qc.backends.backend_factory.BackendFactory.create_backend('b200', device='cuda:0')
    engine  = tneq_qc.core.engine_siamese.EngineSiamese(backend=backend, strategy_mode='balanced')

After `register` the reference's own EngineSiamese / Optimizer / tests drive this
package's CUDA path: the backend is registered with the reference's
BackendFactory (backend_factory.py:91-100) and the strategy with its
StrategyCompiler for modes 'balanced' and 'full' (compiler.py:38-54), where it
wins the cost comparison against GreedyStrategy (5e5, greedy_strategy.py:608).
The classes registered are subclasses of the REFERENCE ABCs, created here so
that isinstance checks on the reference side hold.
"""
from __future__ import annotations


def register(tneq_qc_module=None):
    import importlib

    bi = importlib.import_module("tneq_qc.backends.backend_interface")
    bf = importlib.import_module("tneq_qc.backends.backend_factory")
    importlib.import_module("tneq_qc.core.engine_siamese")      # resolves the reference's import cycle
    cb = importlib.import_module("tneq_qc.contractor.base")
    cc = importlib.import_module("tneq_qc.contractor.compiler")

    from .backends.backend_b200 import B200Backend
    from .contractor.b200_strategy import B200Strategy

    ref_tnt = importlib.import_module("tneq_qc.core.tn_tensor").TNTensor
    RefBackend = type("B200Backend", (B200Backend, bi.ComputeBackend), {})
    RefStrategy = type("B200Strategy", (B200Strategy, cb.ContractionStrategy), {"tntensor_cls": ref_tnt})
    bf.BackendFactory.register_backend("b200", RefBackend)
    cc.StrategyCompiler.register_strategy(RefStrategy(), modes=["balanced", "full"])
    return RefBackend, RefStrategy
