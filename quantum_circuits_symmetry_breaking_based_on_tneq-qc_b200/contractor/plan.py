"""ContractionPlan: everything that can be decided before seeing a batch.

Built once per (graph, operand signature, dtype) and cached by the strategy:
the symbolic greedy schedule (bit-exact reference bookkeeping), the pairwise
contraction graph, and the lowered device programs

    'fwd'    probabilities / amplitudes          (contract_with_compiled_strategy)
    'bwd'    reverse sweep seeded by autograd    (torch.autograd route through compute_fn)
    'train'  forward + fused loss + reverse sweep (contract_with_compiled_strategy_for_gradient)

The reference rebuilds the equivalent bookkeeping on every call
(greedy_strategy.py:45-598).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from . import cgraph, vm_program
from .greedy_plan import GreedySchedule, build_schedule


def signature_of(nqubits, circuit_states, measure_matrices):
    """(state_dims, mx_info) with the reference's presence rules
    (greedy_strategy.py:108-158): containers may be None, dict, list or tuple;
    list entries past the end and None measurements are simply absent."""

    def present(container, q):
        if container is None:
            return False
        if isinstance(container, dict):
            return q in container
        if isinstance(container, (list, tuple)):
            return q < len(container)
        return True

    state_dims, mx_info = {}, {}
    for q in range(nqubits):
        if present(circuit_states, q):
            state_dims[q] = int(circuit_states[q].shape[0])
        if present(measure_matrices, q) and measure_matrices[q] is not None:
            m = measure_matrices[q]
            bsym = "a" if m.ndim == 3 else ("ab" if m.ndim == 4 else "")
            mx_info[q] = (bsym, int(m.shape[-2]), int(m.shape[-1]))
    return state_dims, mx_info


class ContractionPlan:
    def __init__(self, adjacency_table, nqubits: int, core_shapes: Dict[str, Tuple[int, ...]],
                 state_dims: Dict[int, int], mx_info: Dict[int, Tuple[str, int, int]], dtype: str,
                 right="symmetric", right_table=None, right_core_shapes=None,
                 trainable: Optional[Sequence[Tuple[str, object]]] = None):
        """dtype: 'float32' | 'float64' | 'complex64' | 'complex128'."""
        self.nqubits = nqubits
        self.dtype = dtype
        self.complex_mode = dtype.startswith("complex")
        self.real_dtype = "f32" if dtype in ("float32", "complex64") else "f64"
        self.schedule: GreedySchedule = build_schedule(adjacency_table, nqubits, state_dims, mx_info,
                                                       right=right, right_table=right_table)
        self.core_shapes = {("core", k): tuple(v) for k, v in core_shapes.items()}
        for k, v in (right_core_shapes or {}).items():
            self.core_shapes[("rcore", k)] = tuple(v)
        if trainable is None:
            trainable = list(self.core_shapes)
        self.trainable = list(trainable)
        self.nb = 2 if any(b == "ab" for b, _, _ in mx_info.values()) else 1
        self.mx_batch = {q: b for q, (b, _, _) in mx_info.items()}
        self._graphs: Dict[str, cgraph.CGraph] = {}
        self._programs: Dict[str, vm_program.VMProgram] = {}

    @property
    def equations(self) -> List[str]:
        return self.schedule.equations

    def mps_chain_rank(self) -> int:
        """K if this is a single-layer MPS sweep that the register-resident chain kernel covers
        (csrc/tnq_chain.cu): float32, every group of the canonical form, uniform edge rank 2..4,
        plain (B,K,K) measurements on every qubit; else 0."""
        n = self.nqubits
        eq = self.schedule.equations
        if self.dtype != "float32" or self.nb != 1 or n < 2 or n > 64 or len(eq) != n:
            return 0
        want = ["cdef,c,aeg,higj,h,d,i->ajf"] + ["cdef,aeg,higj,ahc,d,i->ajf"] * (n - 2) + ["acd,adc->a"]
        if n == 2:
            want = ["cdef,c,aeg,higj,h,d,i->ajf", "acd,adc->a"]
        if eq != want or len(self.core_shapes) != n - 1 or any(k[0] != "core" for k in self.core_shapes):
            return 0
        ranks = {d for shp in self.core_shapes.values() for d in shp}
        for st in self.schedule.steps:
            ranks |= set(st.dims.values())
        if len(ranks) != 1:
            return 0
        K = ranks.pop()
        return K if 2 <= K <= 4 else 0

    def mps_ladder(self):
        """(K, [A_0..A_{n-2}], [X_0..X_{n-2}]) if this is the two-layer merged MPS sweep
        (QCTN.merge(mps_n, mps_n), reference qctn.py:1296-1506) that the warp-level ladder kernel
        covers (csrc/tnq_ladder.cu): float32, uniform edge rank 2 or 3, circuit states and plain
        (B,K,K) measurements on every qubit, n >= 3; else None.  A_q / X_q are the core names of the
        first / second layer acting on wires (q, q+1), read off the operands of the greedy groups
        (the bit-exact bookkeeping stays the plan compiler's; the kernel only re-associates the
        arithmetic inside a group)."""
        n = self.nqubits
        eq = self.schedule.equations
        if self.dtype != "float32" or self.nb != 1 or n < 3 or n > 64 or len(eq) != n:
            return None
        want = (["cdef,eghi,c,ahj,klmn,mojp,k,d,l->agnpfio"]
                + ["cdef,ghij,aik,lmno,pqkr,aelpcgn,d,m->ahorfjq"] * (n - 3)
                + ["cdef,gfhi,ahj,klmn,onjp,aekocgm,d,l->api", "acd,adc->a"])
        if eq != want or len(self.core_shapes) != 2 * (n - 1) or any(k[0] != "core" for k in self.core_shapes):
            return None
        ranks = {d for shp in self.core_shapes.values() for d in shp}
        for st in self.schedule.steps:
            ranks |= set(st.dims.values())
        if len(ranks) != 1:
            return None
        K = ranks.pop()
        if K not in (2, 3):
            return None
        layer1, layer2 = [], []
        for q in range(n - 1):
            ops = self.schedule.steps[q].operands
            if ops[0].kind != "core" or ops[1].kind != "core":
                return None
            layer1.append(ops[0].key)
            layer2.append(ops[1].key)
        if len(set(layer1 + layer2)) != 2 * (n - 1):
            return None
        return K, layer1, layer2

    def graph(self, mode: str) -> cgraph.CGraph:
        key = "fwd" if mode == "fwd" else mode
        if key not in self._graphs:
            g = cgraph.build_forward(self.schedule, self.complex_mode, self.core_shapes,
                                     trainable=() if mode == "fwd" else self.trainable)
            if mode == "bwd":
                cgraph.add_backward(g, "input")
            elif mode == "train":
                cgraph.add_backward(g, "loss")
            self._graphs[key] = g
        return self._graphs[key]

    def program(self, mode: str) -> vm_program.VMProgram:
        if mode not in self._programs:
            self._programs[mode] = vm_program.lower(self.graph(mode), mode, self.real_dtype, nb=self.nb)
        return self._programs[mode]
