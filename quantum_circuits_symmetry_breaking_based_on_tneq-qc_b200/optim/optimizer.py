"""Training loop + single step (reference: tneq_qc/optim/optimizer.py:5-285).

Same constructor arguments and hooks (summary_writer, eval_every/eval_fn,
save_every/checkpoint_fn attributes).  `optimize` drives
engine.contract_with_compiled_strategy_for_gradient; the reference's version of
that line has a typo (`self.enxgine`, optimizer.py:85, SURVEY defect D3) and
cannot run -- the intended behaviour is implemented here.
"""
from __future__ import annotations

from typing import Optional


class Optimizer:
    def __init__(self, method="adam", learning_rate=0.01, max_iter=1000, tol=1e-6, beta1=0.9, beta2=0.999,
                 epsilon=1e-8, engine=None, lr_schedule: Optional[list] = None, momentum=0.0, stiefel=True,
                 verbose=True):
        self.method, self.learning_rate, self.max_iter, self.tol = method, learning_rate, max_iter, tol
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.momentum, self.stiefel, self.lr_schedule = momentum, stiefel, lr_schedule
        self.engine = engine
        self.iter = 0
        self.opt_state = {}
        self.verbose = verbose

    def _apply_lr_schedule(self):
        """lr_schedule = [(step, lr), ...] ascending; the last entry with step <= iter wins."""
        if self.lr_schedule is None:
            return
        for step, lr in reversed(self.lr_schedule):
            if self.iter >= step:
                self.learning_rate = lr
                return

    def optimize(self, qctn, data_list, **kwargs):
        loss_value = None
        while self.iter < self.max_iter:
            data = data_list[self.iter % len(data_list)]
            loss, grads = self.engine.contract_with_compiled_strategy_for_gradient(qctn, **data, **kwargs)
            loss_value = float(loss) if hasattr(loss, "item") else loss
            self._apply_lr_schedule()
            writer = getattr(self, "summary_writer", None)
            if writer is not None:
                try:
                    writer.add_scalar("train/loss", loss_value, self.iter)
                except Exception:
                    pass
            if self.tol and loss_value < self.tol:
                print(f"Convergence achieved at iteration {self.iter} with loss {loss_value}.")
                break
            if self.verbose:
                print(f"Iteration {self.iter}: loss = {loss_value} lr = {self.learning_rate}")
            self.step(qctn, grads)
            eval_every, eval_fn = getattr(self, "eval_every", 0), getattr(self, "eval_fn", None)
            if eval_every and eval_fn is not None and (self.iter + 1) % eval_every == 0:
                try:
                    metrics = eval_fn(self.iter + 1, qctn)
                except Exception as e:
                    print(f"[Optimizer] Eval function raised an exception at iter {self.iter + 1}: {e}")
                    metrics = None
                if metrics and writer is not None:
                    for name, value in metrics.items():
                        try:
                            writer.add_scalar(f"eval/{name}", float(value), self.iter + 1)
                        except Exception:
                            pass
            save_every, ckpt = getattr(self, "save_every", 0), getattr(self, "checkpoint_fn", None)
            if save_every and ckpt is not None and (self.iter + 1) % save_every == 0:
                try:
                    ckpt(self.iter + 1, qctn, loss_value)
                except Exception as e:
                    print(f"[Optimizer] Checkpoint function raised an exception at iter {self.iter + 1}: {e}")
            self.iter += 1
        else:
            print(f"Maximum iterations reached: {self.max_iter} with final loss {loss_value}.")

    def step(self, qctn, grads):
        """One update of every core, in qctn.cores order (optimizer.py:250-284)."""
        keys = qctn.cores
        params = [qctn.cores_weights[k] for k in keys]
        if len(grads) != len(params):
            raise ValueError(f"{len(grads)} gradients for {len(params)} cores: the engine returns gradients only for cores "
                             "with requires_grad=True (flag every core, as examples/train_single_node.py does)")
        hp = dict(learning_rate=self.learning_rate, beta1=self.beta1, beta2=self.beta2, epsilon=self.epsilon,
                  iter=self.iter, momentum=self.momentum, stiefel=self.stiefel)
        new_params, self.opt_state = self.engine.backend.optimizer_update(params, grads, self.opt_state, self.method, hp)
        for k, p in zip(keys, new_params):
            qctn.cores_weights[k] = p
