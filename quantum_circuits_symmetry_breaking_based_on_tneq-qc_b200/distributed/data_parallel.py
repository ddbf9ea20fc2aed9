"""Data-parallel training over the GPUs of one box (batch sharding, replicated cores).

Mirror of the reference trainer (tneq_qc/distributed/parallel/data_parallel.py:28-424):
same TrainingConfig fields, same method names, same semantics --

    train_step: local forward+backward on this rank's batch
                -> gradients averaged over ranks (equal weight per rank: every rank's loss
                   is the mean over ITS batch, data_parallel.py:266-307)
                -> identical optimizer step on every rank
                -> loss averaged over ranks for logging

with the deviations the SURVEY lists as necessary:
  * cores are broadcast from rank 0 once (`sync_model_weights`); the reference builds an
    independent random QCTN per rank and never synchronises it (distributed_trainer.py:290-343);
  * gradients AND the loss travel in ONE packed NCCL all-reduce per step (NVLink/NVSwitch);
    the reference issues one blocking collective per core plus one for the loss
    (comm_torch.py:510-522) and calls a method that does not exist (defect D9);
  * the SGDG step draws `random.randint` for its 1 % QR retraction (backend_pytorch.py:382): every
    trainer owns a dedicated `random.Random(config.seed)` that the optimizer state carries ('rng'),
    so the replicas stay in lock step whatever else in the process uses the `random` module.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from .comm_nccl import NcclComm, ReduceOp


@dataclass
class TrainingConfig:
    """data_parallel.py:28-52"""
    learning_rate: float = 1e-2
    beta1: float = 0.9
    beta2: float = 0.95
    epsilon: float = 1e-8
    tol: float = 1e-6
    max_steps: int = 1000
    log_interval: int = 10
    checkpoint_interval: int = 100
    checkpoint_dir: Optional[str] = None
    gradient_accumulation_steps: int = 1
    async_gradient_sync: bool = False
    optimizer_method: str = "sgdg"
    momentum: float = 0.9
    stiefel: bool = True
    seed: int = 42
    # (no reference counterpart) large-bond route: start every core's all-reduce when its gradient is final, under the
    # rest of the reverse sweep, instead of one exchange after the step
    overlap_exchange: bool = True


@dataclass
class TrainingStats:
    step: int = 0
    loss: float = float("inf")
    best_loss: float = float("inf")
    total_time: float = 0.0


class DataParallelTrainer:
    def __init__(self, engine, qctn, config: Optional[TrainingConfig] = None, comm: Optional[NcclComm] = None,
                 optimizer=None):
        self.engine, self.qctn = engine, qctn
        self.config = config or TrainingConfig()
        self.mpi = self.comm = comm or NcclComm()
        self.rank, self.world_size = self.comm.rank, self.comm.world_size
        if optimizer is None:
            from ..optim.optimizer import Optimizer
            c = self.config
            optimizer = Optimizer(method=c.optimizer_method, learning_rate=c.learning_rate, max_iter=c.max_steps,
                                  tol=c.tol, beta1=c.beta1, beta2=c.beta2, epsilon=c.epsilon, engine=engine,
                                  momentum=c.momentum, stiefel=c.stiefel, verbose=False)
        self.optimizer = optimizer
        self.global_step = 0
        self.stats = TrainingStats()
        self.accumulated_grads: Optional[List[torch.Tensor]] = None
        self.accumulation_count = 0
        self._oneshot, self._oneshot_tried = None, False
        self._graphs, self._epi_fn, self._ls_dev = False, None, None
        self._ovl_fn, self._ovl = None, None
        # lock-step QR retractions on every rank: a private stream, not the process-global one
        self.optimizer.opt_state["rng"] = random.Random(self.config.seed)

    def _log(self, msg: str, level: str = "info"):
        if self.comm.is_main_process():
            print(f"[{level.upper()}] {msg}")

    # ---- data ------------------------------------------------------------------------
    def partition_data(self, data_list: List[Dict]) -> List[Dict]:
        """Contiguous slices of the LIST of batches; earlier ranks take the remainder
        (data_parallel.py:142-170)."""
        n, w, r = len(data_list), self.world_size, self.rank
        per, rem = divmod(n, w)
        start = r * per + min(r, rem)
        return data_list[start:start + per + (1 if r < rem else 0)]

    def sync_model_weights(self, src: int = 0):
        """Make every replica start from rank `src`'s cores (one packed broadcast)."""
        raws, tnts = [], []
        for name in self.qctn.cores:
            w = self.qctn.cores_weights[name]
            if hasattr(w, "scale") and hasattr(w, "tensor"):
                raws.append(w.tensor)
                tnts.append(w)
            else:
                raws.append(w)
        self.comm.broadcast_tensors_packed(raws, src=src)
        if tnts:                                   # TNTensor cores: the host-side scales travel too
            scales = self.comm.broadcast_object([(float(w.scale), float(w.log_scale)) for w in tnts], src=src)
            for w, (sc, ls) in zip(tnts, scales):
                w.scale, w.log_scale = sc, ls

    # ---- one step ---------------------------------------------------------------------
    def compute_local_gradients(self, data: Dict, circuit_states_list: List) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        loss, grads = self.engine.contract_with_compiled_strategy_for_gradient(
            self.qctn, circuit_states_list=circuit_states_list, **data)
        return loss, list(grads)

    def sync_gradients(self, local_grads: List[torch.Tensor]) -> List[torch.Tensor]:
        return self.comm.allreduce_list(local_grads, op=ReduceOp.AVG)

    def sync_loss(self, local_loss: float) -> float:
        return self.comm.allreduce_scalar(float(local_loss), op=ReduceOp.AVG)

    def _oneshot_for(self, n_grad: int, device) -> Optional[object]:
        """The one-shot NVLink all-reduce (csrc/tnq_allreduce.cu) for small float32 messages on NCCL
        process groups; None when symmetric memory is unavailable (the NCCL path is used then)."""
        if self._oneshot is None and not self._oneshot_tried:
            self._oneshot_tried = True
            if self.comm.backend == "nccl" and self.world_size > 1 and n_grad < (1 << 18):
                from .oneshot import OneShotAllReduce
                self._oneshot = OneShotAllReduce.create(n_grad + 16, device)
        return self._oneshot

    def sync_gradients_and_loss(self, grads: List[torch.Tensor], loss) -> Tuple[List[torch.Tensor], torch.Tensor]:
        """Gradients and loss in ONE exchange: this repository's one-shot NVLink kernel when the
        gradients sit in one flat float32 buffer (the fused CUDA routes return them that way), else one
        packed NCCL all-reduce."""
        loss_t = loss if isinstance(loss, torch.Tensor) else torch.tensor(float(loss), device=grads[0].device)
        loss_t = loss_t.detach().reshape(1).to(grads[0].real.dtype if grads[0].is_complex() else grads[0].dtype)
        base = getattr(grads[0], "_base", None) if len(grads) else None
        n_grad = sum(g.numel() for g in grads)
        if (base is not None and base.is_cuda and base.dtype == torch.float32 and base.is_contiguous()
                and base.numel() == n_grad and all(getattr(g, "_base", None) is base for g in grads)
                and loss_t.dtype == torch.float32):
            red = self._oneshot_for(n_grad, base.device)
            if red is not None and n_grad + 1 <= red.nmax:
                out = red.mean(base, loss_t)
                pieces, at = [], 0
                for g in grads:
                    pieces.append(out[at:at + g.numel()].reshape(g.shape))
                    at += g.numel()
                return pieces, out[-1]
        out = self.comm.allreduce_list(list(grads) + [loss_t], op=ReduceOp.AVG)
        return out[:-1], out[-1][0]

    def _overlap_for(self, fn) -> bool:
        """Large-bond route (GB-sized gradients, e.g. 15 cores x 134 MB at bond 64): every core's NCCL all-reduce is
        started on a side stream by compute_fn.set_grad_ready_hook as soon as that core's gradient is final -- the
        reverse sweep finishes the cores in reverse-use order -- and overlaps the rest of the sweep.  The reference
        issues one blocking collective per core after the backward pass (comm_torch.py:292-318, 510-522)."""
        if not (self.config.overlap_exchange and self.world_size > 1 and hasattr(fn, "set_grad_ready_hook")
                and self.comm.device.type == "cuda"):
            return False
        if self._ovl_fn is not fn:
            import torch.distributed as dist
            stream = torch.cuda.Stream(device=self.comm.device)
            state = {"fired": 0, "stream": stream}
            world = self.world_size

            def ready(name, flat):
                stream.wait_stream(torch.cuda.current_stream(flat.device))
                with torch.cuda.stream(stream):
                    dist.all_reduce(flat)
                    flat.div_(world)
                state["fired"] += 1

            fn.set_grad_ready_hook(ready)
            self._ovl_fn, self._ovl = fn, state
        self._ovl["fired"] = 0
        return True

    def enable_cuda_graphs(self, flag: bool = True):
        """One CUDA-graph launch per training step (no reference counterpart): the fused contraction kernels
        AND the one-shot NVLink exchange of gradients + loss are captured together (the engine's graph replay,
        EngineSiamese.enable_cuda_graphs, with the exchange as its epilogue).  Same contract as the engine's
        replay: cores updated in place (the 'sgdg' flat step does that), batches copied into static buffers."""
        self._graphs = bool(flag)
        self.engine.enable_cuda_graphs(flag)
        self.optimizer.opt_state["pingpong"] = bool(flag)     # optim/steps.py: cores alternate between two buffers
        if not flag and self._epi_fn is not None:
            self._epi_fn.set_graph_epilogue(None)
            self._epi_fn = None

    def _graph_epilogue_for(self, fn, grads_like_numel: int, device):
        """Register the exchange as the epilogue of fn's training graph (once per compiled function)."""
        if fn is self._epi_fn:
            return
        if self._epi_fn is not None:
            self._epi_fn.set_graph_epilogue(None)
        self._epi_fn = None
        red = self._oneshot_for(grads_like_numel, device)
        if red is None or not hasattr(fn, "set_graph_epilogue"):
            return
        self._ls_dev = torch.zeros(1, dtype=torch.float32, device=device)

        def epilogue(loss0, grads):
            base = getattr(grads[0], "_base", None)
            if base is None or base.dtype != torch.float32 or base.numel() != grads_like_numel:
                return None
            # the captured kernel computes the loss with log_scale = 0; the step's log_scale arrives in _ls_dev
            return red.mean(base, loss0.reshape(1) - self._ls_dev)

        fn.set_graph_epilogue(epilogue)
        self._epi_fn = fn

    def accumulate_gradients(self, grads: List[torch.Tensor]):
        if self.accumulated_grads is None:
            self.accumulated_grads = [g.clone() for g in grads]
        else:
            for a, g in zip(self.accumulated_grads, grads):
                a += g
        self.accumulation_count += 1

    def get_accumulated_gradients(self) -> List[torch.Tensor]:
        if self.accumulated_grads is None:
            raise ValueError("No gradients accumulated")
        avg = [g / self.accumulation_count for g in self.accumulated_grads]
        self.accumulated_grads, self.accumulation_count = None, 0
        return avg

    def train_step(self, data: Dict, circuit_states_list: List) -> float:
        k = self.config.gradient_accumulation_steps
        fn = None
        if self._graphs and k == 1 and self.world_size > 1 and "measure_input_list" in data:
            mi = data["measure_input_list"]
            fn = self.engine._compiled(self.qctn, circuit_states_list, mi, data.get("measure_is_matrix", True), "symmetric")
            n_grad = sum((w.tensor if hasattr(w, "scale") else w).numel() for w in self.qctn.cores_weights.values())
            self._graph_epilogue_for(fn, n_grad, self.comm.device)
            if self._epi_fn is fn:
                ls = sum(float(getattr(m, "log_scale", 0.0) or 0.0) for m in (mi.values() if isinstance(mi, dict) else mi)
                         if m is not None)
                ls += sum(float(w.log_scale) for w in self.qctn.cores_weights.values() if hasattr(w, "log_scale"))
                self._ls_dev.fill_(ls)
        overlapping = False
        if (k == 1 and self.world_size > 1 and "measure_input_list" in data and self.config.overlap_exchange
                and self.comm.device.type == "cuda" and hasattr(self.engine, "_compiled")):
            ofn = fn if fn is not None else self.engine._compiled(
                self.qctn, circuit_states_list, data["measure_input_list"], data.get("measure_is_matrix", True), "symmetric")
            overlapping = self._overlap_for(ofn)
        loss, grads = self.compute_local_gradients(data, circuit_states_list)
        if overlapping and self._ovl["fired"] == len(grads) and len(grads):
            # every core's gradient was averaged in place while the sweep was still running: wait for the last one
            torch.cuda.current_stream(self.comm.device).wait_stream(self._ovl["stream"])
            self.optimizer.step(self.qctn, list(grads))
            loss_avg = self.sync_loss(float(loss))
            self.optimizer.iter += 1
            return loss_avg
        if fn is not None and self._epi_fn is fn and fn.graph_stats["last_extra"] is not None:
            out = fn.graph_stats["last_extra"]             # exchange done inside the step's graph
            pieces, at = [], 0
            for g in grads:
                pieces.append(out[at:at + g.numel()].reshape(g.shape))
                at += g.numel()
            self.optimizer.step(self.qctn, pieces)
            loss_avg = float(out[-1])
            if loss_avg != loss_avg and self._oneshot is not None:
                self._oneshot.check()
            self.optimizer.iter += 1
            return loss_avg
        if k > 1:
            self.accumulate_gradients(grads)
            if (self.global_step + 1) % k == 0:
                g, _ = self.sync_gradients_and_loss(self.get_accumulated_gradients(), loss)
                self.optimizer.step(self.qctn, g)
            loss_avg = self.sync_loss(float(loss))
        else:
            g, loss_avg = self.sync_gradients_and_loss(grads, loss)
            self.optimizer.step(self.qctn, g)
            loss_avg = float(loss_avg)
            if loss_avg != loss_avg and self._oneshot is not None:
                self._oneshot.check()              # NaN: did the exchange give up on a missing peer?
        self.optimizer.iter += 1
        return loss_avg

    def train(self, data_list: List[Dict], circuit_states_list: List, max_steps: Optional[int] = None) -> TrainingStats:
        """Every rank walks ITS partition of data_list round-robin (data_parallel.py:311-387)."""
        import time
        local = self.partition_data(data_list)
        if not local:
            raise ValueError(f"rank {self.rank} received no data: {len(data_list)} batches for {self.world_size} ranks")
        self.sync_model_weights()
        steps = max_steps or self.config.max_steps
        t0 = time.time()
        for _ in range(steps):
            loss = self.train_step(local[self.global_step % len(local)], circuit_states_list)
            self.global_step += 1
            self.stats.step, self.stats.loss = self.global_step, loss
            self.stats.best_loss = min(self.stats.best_loss, loss)
            if self.config.log_interval and self.global_step % self.config.log_interval == 0:
                self._log(f"step {self.global_step}: loss = {loss:.6f}")
            if self.config.tol and loss < self.config.tol:
                break
        self.stats.total_time = time.time() - t0
        return self.stats
