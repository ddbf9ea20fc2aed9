"""B200Backend: the ComputeBackend whose tensors live on a B200 and whose
contraction hot path is hand-written sm_100a CUDA (libtneq_b200.so).

Interface: tneq_qc/backends/backend_interface.py:48-518; behaviour of every
method follows the reference's PyTorch backend
(tneq_qc/backends/backend_pytorch.py:13-664) so the engine, the optimizer and
the tests can switch backends by name:

    BackendFactory.register_backend('b200', B200Backend)
    backend = BackendFactory.create_backend('b200', device='cuda:0', dtype='float32')

PyTorch is used here for device memory, streams, autograd plumbing and the
small element-wise helpers; there is NO CPU fallback: constructing the backend
without a CUDA device, or without the built library, raises.
"""
from __future__ import annotations

import random
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from .backend_interface import BackendInfo, ComputeBackend

_DTYPES = {"float32": torch.float32, "float64": torch.float64, "complex64": torch.complex64,
           "complex128": torch.complex128, "complex": torch.complex64}


class B200Backend(ComputeBackend):
    def __init__(self, device: Optional[str] = None, dtype: Optional[Any] = None, tensor_type: Optional[str] = None):
        super().__init__(tensor_type=tensor_type)
        self.torch = torch
        if device is None:
            device = "cuda"
        if not str(device).startswith("cuda"):
            raise RuntimeError(f"B200Backend runs on CUDA devices only (got device={device!r}); "
                               "tneq_b200 has no CPU fallback")
        if not torch.cuda.is_available():
            raise RuntimeError("B200Backend: no CUDA device is visible; tneq_b200 has no CPU fallback")
        _lib.load()  # fail loudly now if the extension has not been built
        if dtype is None:
            self.default_dtype = torch.float32
        elif isinstance(dtype, str):
            if dtype not in _DTYPES:
                raise ValueError(f"Unsupported dtype string '{dtype}' for B200Backend. Supported: {list(_DTYPES)}")
            self.default_dtype = _DTYPES[dtype]
        else:
            self.default_dtype = dtype
        self._device = torch.device(device)
        name = {v: k for k, v in _DTYPES.items() if k != "complex"}.get(self.default_dtype, str(self.default_dtype))
        self.backend_info = BackendInfo("b200", device=str(device), dtype=name)

    # -- identity -------------------------------------------------------------
    def get_backend_name(self) -> str:
        return "b200"

    def _get_raw_tensor_type(self):
        return torch.Tensor

    def set_random_seed(self, seed: int):
        torch.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
        np.random.seed(seed)
        random.seed(seed)

    # -- conversion -------------------------------------------------------------
    def convert_to_tensor(self, array):
        if isinstance(array, torch.Tensor):
            t = array
            if t.device != self._device and not (t.is_cuda and self._device.index is None):
                t = t.to(self._device)
            return t if t.dtype == self.default_dtype else t.to(self.default_dtype)
        if not isinstance(array, np.ndarray):
            array = np.array(array)
        return torch.as_tensor(array, dtype=self.default_dtype).to(self._device)

    def tensor_to_numpy(self, tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor)
        return tensor.detach().cpu().numpy()

    # -- execution ----------------------------------------------------------------
    def execute_expression(self, expression, *tensors):
        return expression(*tensors)

    def jit_compile(self, func):
        return func

    def compute_value_and_grad(self, loss_fn, argnums):
        """(loss, grads) of loss_fn w.r.t. the positional args listed in argnums
        (backend_pytorch.py:107-166).  The reverse sweep of the contraction runs as
        one CUDA program inside compute_fn's autograd node."""
        idx = list(argnums) if isinstance(argnums, (list, tuple, range)) else [argnums]

        def value_and_grad_fn(*args):
            leaves = []
            for i, a in enumerate(args):
                t = self.convert_to_tensor(a)
                if i in idx:
                    if not t.is_leaf:
                        t = t.detach()
                    t.requires_grad_(True)
                else:
                    t = t.detach()
                leaves.append(t)
            loss = loss_fn(*leaves)
            target = loss.real if torch.is_complex(loss) else loss
            if target.ndim > 0:
                target = target.sum()
            grads = torch.autograd.grad(target, [leaves[i] for i in idx], create_graph=False, retain_graph=False)
            out = (loss.real if torch.is_complex(loss) else loss).detach()
            return (out.sum() if out.ndim > 0 else out), grads

        return value_and_grad_fn

    # -- optimizers (backend_pytorch.py:200-468) ------------------------------------
    def optimizer_update(self, params: List[Any], grads: List[Any], state: Dict[str, Any], method: str,
                         hyperparams: Dict[str, Any]) -> Tuple[List[Any], Dict[str, Any]]:
        from ..optim import steps
        return steps.optimizer_update(params, grads, state, method, hyperparams)

    def init_random_core(self, shape):
        """QR-orthogonal matrix reshaped to `shape` (backend_pytorch.py:470-495)."""
        d = int(np.prod(shape[: len(shape) // 2]))
        m = torch.randn((d, d), device=self._device, dtype=self.default_dtype)
        q, r = torch.linalg.qr(m)
        diag = torch.diag(r)
        if torch.is_complex(diag):
            q = q @ torch.diag((diag / (diag.abs() + 1e-12)).conj())
        else:
            q = q * torch.sign(diag).unsqueeze(0)
        return self.wrap_tensor(q.reshape(shape))

    # -- small tensor helpers ----------------------------------------------------------
    def reshape(self, tensor, shape):
        return tensor.reshape(shape)

    def eye(self, n: int, dtype=None):
        return torch.eye(n, dtype=dtype or self.default_dtype, device=self._device)

    def zeros(self, shape, dtype=None):
        return torch.zeros(shape, dtype=dtype or self.default_dtype, device=self._device)

    def ones(self, shape, dtype=None):
        return torch.ones(shape, dtype=dtype or self.default_dtype, device=self._device)

    def clone(self, tensor):
        return tensor.clone()

    def unsqueeze(self, tensor, dim):
        return tensor.unsqueeze(dim)

    def expand(self, tensor, *sizes):
        return tensor.expand(*sizes)

    def clamp(self, tensor, min=None, max=None):
        if torch.is_complex(tensor):  # real part only, as the reference does
            return torch.complex(torch.clamp(tensor.real, min=min, max=max), tensor.imag)
        return torch.clamp(tensor, min=min, max=max)

    def diagonal(self, tensor, dim1=-2, dim2=-1):
        return torch.diagonal(tensor, dim1=dim1, dim2=dim2)

    def sum(self, tensor, dim=None, keepdim=False):
        return torch.sum(tensor) if dim is None else torch.sum(tensor, dim=dim, keepdim=keepdim)

    def multinomial(self, probs, num_samples):
        return torch.multinomial(probs, num_samples=num_samples)

    def arange(self, *args, dtype=None):
        return torch.arange(*args, dtype=dtype or torch.long, device=self._device)

    def stack(self, tensors, dim=0):
        return torch.stack(tensors, dim=dim)

    def log(self, tensor):
        return torch.log(tensor)

    def mean(self, tensor, dim=None, keepdim=False):
        return torch.mean(tensor) if dim is None else torch.mean(tensor, dim=dim, keepdim=keepdim)

    def squeeze(self, tensor, dim=None):
        return tensor.squeeze() if dim is None else tensor.squeeze(dim)

    def einsum(self, equation, *operands):
        return torch.einsum(equation, *operands)

    def detach(self, tensor):
        return tensor.detach() if hasattr(tensor, "detach") else tensor

    def lgamma(self, tensor):
        return torch.lgamma(tensor)

    def exp(self, tensor):
        return torch.exp(tensor)

    def sqrt(self, tensor):
        return torch.sqrt(tensor)

    def square(self, tensor):
        return torch.square(tensor)

    def permute(self, tensor, dims):
        return tensor.permute(dims)

    def ones_like(self, tensor):
        return torch.ones_like(tensor)

    def linspace(self, start, end, steps, dtype=None):
        return torch.linspace(start, end, steps, dtype=dtype or self.default_dtype, device=self._device)

    def cumsum(self, tensor, dim, dtype=None):
        return torch.cumsum(tensor, dim=dim, dtype=dtype)

    def rand(self, size, dtype=None):
        return torch.rand(size, dtype=dtype or self.default_dtype, device=self._device)

    def real(self, tensor):
        return torch.real(tensor)

    def is_complex(self, tensor) -> bool:
        return torch.is_complex(tensor)

    def abs_square(self, tensor):
        if torch.is_complex(tensor):
            return tensor.real * tensor.real + tensor.imag * tensor.imag
        return tensor

    def gather(self, input, dim, index):
        return torch.gather(input, dim, index)
