"""QCTN data model: graph string -> cores + adjacency table.

Host-side mirror of the part of the reference data model that the contraction
plan compiler consumes (reference: tneq_qc/core/qctn.py).  Same public names,
argument meaning and on-disk checkpoint format, written from scratch:

  QCTNHelper.generate_example_graph   qctn.py:34-425   (mps / tree / wall strings)
  QCTN.__init__ / adjacency_table     qctn.py:482-536, 591-714
  QCTN._init_cores                    qctn.py:724-757
  QCTN.set_cores                      qctn.py:762-900
  QCTN.save_cores / load_cores        qctn.py:902-984  (safetensors, core_<sym>[_real|_imag])
  QCTN.split / merge / merge_with     qctn.py:1296-1523

A graph is one text line per qubit, e.g. ``-2-a-4-b-2-``: numbers are edge
ranks, single non-digit non-dash characters are cores.  Cores are ordered by
``symbol_of`` (the opt_einsum symbol order: a-z, A-Z, then chr(i+140)).
adjacency_table[i] = {core_idx, core_name, in_edge_list, out_edge_list,
input_shape, output_shape, input_dim, output_dim}; every edge is
{neighbor_idx, neighbor_name, edge_rank, qubit_idx} with neighbor_idx == -1 for
circuit inputs / outputs.  Core tensors have shape input_shape + output_shape.
"""
from __future__ import annotations

import re
import warnings
from pathlib import Path
from typing import Mapping, Optional, Union

import numpy as np

from .tn_tensor import TNTensor

_ALPHABET = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def symbol_of(i: int) -> str:
    """Core / index symbol number i (same sequence as opt_einsum.get_symbol)."""
    if i < 52:
        return _ALPHABET[i]
    return chr(i + 2048) if i >= 55296 else chr(i + 140)


_SYMBOL_RANK = {symbol_of(i): i for i in range(10000)}


# ----------------------------------------------------------------------------
# example graph strings
# ----------------------------------------------------------------------------
def _mps_lines(n, d):
    """Staircase of n-1 two-qubit cores; core j sits on qubits j, j+1."""
    sym = [symbol_of(i) for i in range(n)]
    gap = "-" * 6
    rows = []
    for q in range(n):
        if q == 0:
            body = sym[0] + gap * (n - 2)
        elif q == n - 1:
            body = gap * (n - 2) + sym[q - 1]
        else:
            body = gap * (q - 1) + sym[q - 1] + f"--{d}--" + sym[q] + gap * (n - q - 2)
        rows.append(f"-{d}-{body}-{d}-\n")
    return "".join(rows)


def _tree_lines(n, d):
    """Two staircases meeting in the middle (qctn.py:72-134)."""
    sym = [symbol_of(i) for i in range(n)]
    half = n // 2
    rows = []
    for i in range(half):
        pad_l = "-" * (4 * (half - 1 - i))
        if i == 0:
            body = pad_l + sym[0]
        else:
            body = pad_l + sym[i] + f"-{d}-" + sym[i - 1] + "-" * (4 * (i - 1))
        rows.append(body)
    if n % 2 == 1:
        rows.append(sym[half - 1] + "-" * (4 * (half - 1)))
    for i in range(half, 2 * half):
        pad_l = "-" * (4 * (i - half))
        if i < 2 * half - 1:
            body = pad_l + sym[i - 1] + f"-{d}-" + sym[i] + "-" * (4 * (2 * half - 2 - i))
        else:
            body = pad_l + sym[i - 1]
        rows.append(body)
    return "".join(f"-{d}-{b}-{d}-\n" for b in rows)


def _wall_lines(n, layers, d):
    """Brick wall of two-qubit cores, `layers` columns (qctn.py:232-288)."""
    cells = [["-"] * (4 * layers) for _ in range(n)]
    for row in cells:
        row[-2] = d
    per_row = layers // 2
    k = 0
    for q in range(n - 1):
        shift = 0 if q % 2 == 0 else 4
        for j in range(per_row):
            col = shift + 8 * j
            name = symbol_of(k)
            k += 1
            cells[q][col] = name
            cells[q + 1][col] = name
            last = j == per_row - 1
            if not last or q > 0:
                cells[q][col + 2] = d
            if not last or q != n - 2:
                cells[q + 1][col + 2] = d
    return "\n".join(f"-{d}-" + "".join(r) for r in cells)


class QCTNHelper:
    """Graph-string generators (reference: qctn.py:11-456)."""

    @staticmethod
    def generate_example_graph(n=16, target=False, graph_type="any", dim_char=None):
        if target:
            return ("-2-A-5-----C-3-----E-2-\n"
                    "-2-----B----4------E-2-\n"
                    "-2-A-4-B-7-C-2-D-4-E-2-\n"
                    "-2-----B-6-----D-----2-\n"
                    "-2-A-3-----C-8-D-----2-")
        d = "3" if dim_char is None else str(dim_char)
        if graph_type == "tree":
            return _tree_lines(n, d)
        if graph_type == "wall":
            return _wall_lines(n, 4, d)
        return _mps_lines(n, d)


# ----------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------
def _tokens(line):
    """'-2-a-5-b-3-' -> [('dim', 2), ('core', 'a'), ('dim', 5), ('core', 'b'), ('dim', 3)]"""
    out = []
    for m in re.finditer(r"\d+|[^\d-]", line.strip()):
        t = m.group()
        out.append(("dim", int(t)) if t[0].isdigit() else ("core", t))
    return out


def _line_of(tokens):
    return "-" + "-".join(str(v) for _, v in tokens) + "-"


class QCTN:
    """Quantum-circuit tensor network built from a graph string."""

    def __init__(self, graph: str, backend=None):
        self.graph = graph
        self.qubits = graph.strip().splitlines()
        self.nqubits = len(self.qubits)
        self.qubit_indices = list(range(self.nqubits))
        self.cores = sorted({c for c in graph if c in _SYMBOL_RANK}, key=_SYMBOL_RANK.__getitem__)
        self.ncores = len(self.cores)
        self.dict_core2idx = {c: i for i, c in enumerate(self.cores)}
        self.backend = backend
        self._loaded_metadata: Optional[Mapping[str, str]] = None
        self._build_adjacency()
        self.cores_weights = {}
        if backend is not None:
            self._init_cores()

    # -- graph -> table ------------------------------------------------------
    def _build_adjacency(self):
        tab = [dict(core_idx=i, core_name=c, in_edge_list=[], out_edge_list=[]) for i, c in enumerate(self.cores)]

        def edge(nbr, rank, q):
            return dict(neighbor_idx=nbr, neighbor_name=self.cores[nbr] if nbr >= 0 else "", edge_rank=rank, qubit_idx=q)

        for q, line in enumerate(self.qubits):
            toks = _tokens(line)
            if len(toks) < 3 or toks[0][0] != "dim" or toks[-1][0] != "dim":
                raise ValueError(f"malformed qubit line {q}: {line!r}")
            chain = [self.dict_core2idx[v] for k, v in toks if k == "core"]
            ranks = [v for k, v in toks if k == "dim"]
            if len(ranks) != len(chain) + 1:
                raise ValueError(f"malformed qubit line {q}: {line!r}")
            # the reference records the circuit input and output edge of a line
            # before the core-to-core links of that line
            tab[chain[0]]["in_edge_list"].append(edge(-1, ranks[0], q))
            tab[chain[-1]]["out_edge_list"].append(edge(-1, ranks[-1], q))
            for j in range(len(chain) - 1):
                tab[chain[j]]["out_edge_list"].append(edge(chain[j + 1], ranks[j + 1], q))
                tab[chain[j + 1]]["in_edge_list"].append(edge(chain[j], ranks[j + 1], q))
        for t in tab:
            t["input_shape"] = [e["edge_rank"] for e in t["in_edge_list"]]
            t["output_shape"] = [e["edge_rank"] for e in t["out_edge_list"]]
            t["input_dim"] = int(np.prod(t["input_shape"])) if t["input_shape"] else 1
            t["output_dim"] = int(np.prod(t["output_shape"])) if t["output_shape"] else 1
        self.adjacency_table = tab

    def core_shape(self, name):
        t = self.adjacency_table[self.dict_core2idx[name]]
        return t["input_shape"] + t["output_shape"]

    def _init_cores(self):
        for t in self.adjacency_table:
            core = self.backend.init_random_core([t["input_dim"], t["output_dim"]])
            self.cores_weights[t["core_name"]] = self.backend.reshape(core, t["input_shape"] + t["output_shape"])

    def __repr__(self):
        return f"QCTN(nqubits={self.nqubits}, ncores={self.ncores})\n{self.graph}"

    # -- setting cores ---------------------------------------------------------
    def _set_single_core(self, name, tensor):
        want = tuple(self.cores_weights[name].shape) if name in self.cores_weights else tuple(self.core_shape(name))
        got = tuple(tensor.shape)
        if int(np.prod(got)) != int(np.prod(want)):
            raise ValueError(f"Core '{name}': size mismatch — input has {int(np.prod(got))} elements "
                             f"(shape {got}) but target has {int(np.prod(want))} elements (shape {want}).")
        if got != want:
            tensor = tensor.reshape(list(want))
        self.cores_weights[name] = tensor

    def set_cores(self, cores, strict: bool = True):
        """Set cores from a list (by position in self.cores) or a dict (by name)."""
        if isinstance(cores, list):
            if strict and len(cores) != self.ncores:
                raise ValueError(f"strict=True: expected {self.ncores} core tensors, got {len(cores)}.")
            if len(cores) != self.ncores:
                warnings.warn(f"strict=False: input list has {len(cores)} tensors but QCTN has {self.ncores} cores.",
                              stacklevel=2)
            for name, t in zip(self.cores, cores):
                self._set_single_core(name, t)
        elif isinstance(cores, dict):
            have, want = set(cores), set(self.cores)
            if strict and have != want:
                bits = []
                if want - have:
                    bits.append(f"missing keys ({len(want - have)}): {want - have}")
                if have - want:
                    bits.append(f"extra keys ({len(have - want)}): {have - want}")
                raise ValueError(f"strict=True: key mismatch — {'; '.join(bits)}.")
            if have != want:
                warnings.warn(f"strict=False: missing {want - have}, ignored {have - want}", stacklevel=2)
            for name in self.cores:
                if name in cores:
                    self._set_single_core(name, cores[name])
        else:
            raise TypeError(f"cores must be a list or dict, got {type(cores).__name__}")

    # -- checkpoints (same key layout as the reference) --------------------------
    def save_cores(self, file_path: Union[str, Path], metadata: Optional[Mapping[str, str]] = None):
        if self.backend is None:
            raise RuntimeError("Backend must be initialized before saving cores.")
        from safetensors.numpy import save_file
        blob = {}
        for name, t in self.cores_weights.items():
            arr = self.backend.tensor_to_numpy(t.tensor * t.scale if isinstance(t, TNTensor) else t)
            if np.iscomplexobj(arr):
                blob[f"core_{name}_real"] = np.ascontiguousarray(arr.real)
                blob[f"core_{name}_imag"] = np.ascontiguousarray(arr.imag)
            else:
                blob[f"core_{name}"] = np.ascontiguousarray(arr)
        save_file(blob, str(file_path), metadata={str(k): str(v) for k, v in (metadata or {}).items()})

    def load_cores(self, file_path: Union[str, Path], strict: bool = True) -> Mapping[str, str]:
        if self.backend is None:
            raise RuntimeError("Backend must be initialized before loading cores.")
        from safetensors import safe_open
        with safe_open(str(file_path), framework="numpy") as f:
            meta = dict(f.metadata() or {})
            keys = set(f.keys())
            for name in self.cores:
                if f"core_{name}_real" in keys:
                    arr = f.get_tensor(f"core_{name}_real") + 1j * f.get_tensor(f"core_{name}_imag")
                elif f"core_{name}" in keys:
                    arr = f.get_tensor(f"core_{name}")
                elif strict:
                    raise KeyError(f"Missing tensor for core {name} in {file_path}")
                else:
                    continue
                t = TNTensor(self.backend.convert_to_tensor(arr))
                t.auto_scale()
                self.cores_weights[name] = t
        self._loaded_metadata = {str(k): str(v) for k, v in meta.items()}
        return self._loaded_metadata

    @classmethod
    def from_pretrained(cls, graph, file_path, backend=None, strict: bool = True) -> "QCTN":
        if backend is None:
            from ..backends.backend_factory import BackendFactory
            backend = BackendFactory.get_default_backend()
        inst = cls(graph, backend=backend)
        inst.load_cores(file_path, strict=strict)
        return inst

    # -- split / merge -------------------------------------------------------------
    _parse_qubit_line = staticmethod(_tokens)
    _rebuild_qubit_line = staticmethod(_line_of)

    @staticmethod
    def _remap_graph(graph_lines, core_map):
        return ["".join(core_map.get(ch, ch) for ch in line) for line in graph_lines]

    def split(self, split_idx=None):
        """Cut into (cores[:split_idx], cores[split_idx:]); boundary bonds become
        outputs of the first and inputs of the second network."""
        if split_idx is None:
            split_idx = self.ncores // 2
        if not 0 < split_idx < self.ncores:
            raise ValueError(f"split_idx must be between 1 and {self.ncores - 1}, got {split_idx}")
        first = set(self.cores[:split_idx])
        lines1, lines2 = [], []
        for q, line in enumerate(self.qubits):
            toks = _tokens(line)
            at1 = [i for i, (k, v) in enumerate(toks) if k == "core" and v in first]
            at2 = [i for i, (k, v) in enumerate(toks) if k == "core" and v not in first]
            if at1 and at2:
                if max(at1) >= min(at2):
                    raise ValueError(f"Cannot split: cores from both groups are interleaved on qubit {q}. "
                                     f"Ensure that all Group-1 cores appear before Group-2 cores on every qubit line.")
                lines1.append(_line_of(toks[: max(at1) + 2]))
                lines2.append(_line_of(toks[min(at2) - 1:]))
            elif at1:
                lines1.append(_line_of(toks))
            elif at2:
                lines2.append(_line_of(toks))
        if not lines1:
            raise ValueError("After split, Group 1 has no qubit lines. All qubits belong to Group 2.")
        if not lines2:
            raise ValueError("After split, Group 2 has no qubit lines. All qubits belong to Group 1.")
        left, right = QCTN("\n".join(lines1), backend=self.backend), QCTN("\n".join(lines2), backend=self.backend)
        for name in self.cores:
            if name in self.cores_weights:
                (left if name in first else right).cores_weights[name] = self.cores_weights[name]
        return left, right

    @staticmethod
    def merge(qctn1, qctn2):
        """Concatenate qctn1 (left) and qctn2 (right) line by line; the shared
        boundary rank is kept once; cores are renamed contiguously."""
        n1, n2 = qctn1.nqubits, qctn2.nqubits
        names = [symbol_of(i) for i in range(qctn1.ncores + qctn2.ncores)]
        map1 = dict(zip(qctn1.cores, names))
        map2 = dict(zip(qctn2.cores, names[qctn1.ncores:]))
        g1, g2 = QCTN._remap_graph(qctn1.qubits, map1), QCTN._remap_graph(qctn2.qubits, map2)
        pad1 = "-" * (max(map(len, g1)) - 3)
        pad2 = "-" * (max(map(len, g2)) - 3)
        rows = []
        for q in range(max(n1, n2)):
            if q < n1 and q < n2:
                m2 = re.match(r"^-\d+-", g2[q])
                rows.append(g1[q] + g2[q][m2.end():])
            elif q < n1:
                m1 = re.search(r"-\d+-$", g1[q])
                rows.append(g1[q][: m1.start()] + pad2 + m1.group())
            else:
                m2 = re.match(r"^-\d+-", g2[q])
                rows.append(m2.group() + pad1 + g2[q][m2.end():])
        merged = QCTN("\n".join(rows), backend=qctn1.backend if qctn1.backend is not None else qctn2.backend)
        for old, new in map1.items():
            if old in qctn1.cores_weights:
                merged.cores_weights[new] = qctn1.cores_weights[old]
        for old, new in map2.items():
            if old in qctn2.cores_weights:
                merged.cores_weights[new] = qctn2.cores_weights[old]
        return merged

    def merge_with(self, other):
        return QCTN.merge(self, other)
