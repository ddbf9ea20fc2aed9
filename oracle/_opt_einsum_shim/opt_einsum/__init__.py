"""Minimal stand-in for the `opt_einsum` package (TEST INFRASTRUCTURE ONLY).

The reference (`/root/reference/tneq_qc`) hard-imports opt_einsum, which is not
installed in this image and cannot be fetched (no network).  The greedy hot
path only ever calls ``opt_einsum.get_symbol`` (tneq_qc/contractor/
greedy_strategy.py:414,899; tneq_qc/core/qctn.py:48,498), which contributes
symbol *names* only, never arithmetic.  This stand-in restates the published
behaviour of ``opt_einsum.parser.get_symbol`` (opt_einsum 3.x, un-pinned by the
reference: it ships no requirements file):

    get_symbol(i) = "abc...zABC...Z"[i]      for i < 52
                  = chr(i + 140)             for 52 <= i < 55296
                  = chr(i + 2048)            otherwise (skips surrogates)

It must be put on sys.path only AFTER ``import torch`` so that
``torch.backends.opt_einsum.is_available()`` stays False and ``torch.einsum``
contracts operands strictly left to right (torch/functional.py), which is the
oracle convention of this repo.
"""

_BASE = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def get_symbol(i: int) -> str:
    if i < 52:
        return _BASE[i]
    if i >= 55296:
        return chr(i + 2048)
    return chr(i + 140)


__all__ = ["get_symbol"]
