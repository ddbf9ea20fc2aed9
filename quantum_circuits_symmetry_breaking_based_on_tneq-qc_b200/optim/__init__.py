from .optimizer import Optimizer

__all__ = ["Optimizer"]
