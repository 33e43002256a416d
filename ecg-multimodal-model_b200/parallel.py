"""Data parallelism for the fusion model: one process per GPU, replicated parameters, the batch
sharded by the caller, gradients averaged by bucketed all-reduce overlapped with backward.

The reference is single-device (config.py:46); semantics follow SURVEY.md section 8e: BatchNorm
stays per-rank (no SyncBN), dropout streams are per-rank, gradients are the mean over ranks.

Every stage of ECGMultimodalModel writes its parameter gradients into one flat fp32 GradArena
laid out in reverse execution order and reports prefixes of it as they become final
(model._Stage._notify).  Each report becomes one asynchronous all-reduce (NCCL: average; run on
NCCL's own stream, which orders itself after the kernels launched so far), so the image
encoder's layer4/layer3/layer2 gradients travel over NVLink while layer1 and the stem are still
in backward.  A callback queued on the autograd engine waits for the outstanding collectives at
the end of backward(), so optimizer.step() may follow immediately, as in train.py:80-81.

The in-place all-reduce of the arena is only the gradient when autograd ADOPTS the arena slices as p.grad, i.e. when
p.grad was None (optimizer.zero_grad() default).  When a stage's parameters already carry gradients -- accumulation
under no_sync(), zero_grad(set_to_none=False) -- autograd adds the arena into p.grad instead, so for that stage the
wrapper does what torch DDP does: it leaves the arena alone during backward and all-reduces the ACCUMULATED p.grad
tensors in the end-of-backward callback (one flat collective, no overlap: the slow but correct path).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DataParallel(torch.nn.Module):
    def __init__(self, module, process_group=None, broadcast=True, min_bucket_elems=0):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # smallest prefix worth a collective of its own (0: every stage report); ECGMM_DP_MIN_BUCKET overrides the default
        self.min_bucket = int(min_bucket_elems) or int(os.environ.get("ECGMM_DP_MIN_BUCKET", "0"))
        self._pending = []
        self._deferred = []  # parameters whose accumulated .grad is reduced at the end of backward
        self._cb_queued = False
        self._enabled = True
        self.buckets_last_step = 0
        self.bytes_last_step = 0
        if broadcast and self.world > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0, group=process_group)
        for st in module.stages():
            object.__setattr__(st, "_grad_ready_cb", lambda arena, upto, _st=st: self._on_ready(arena, upto, _st))

    def forward(self, *a, **k):
        self._pending.clear()
        self._deferred.clear()
        self._cb_queued = False
        self.buckets_last_step = 0
        self.bytes_last_step = 0
        return self.module(*a, **k)

    # ---- called by the stages during backward
    def _on_ready(self, arena, upto, stage=None):
        if self.world <= 1 or not self._enabled:
            return
        mode = getattr(arena, "_dp_deferred", None)  # decided once per arena (arenas are per-backward objects)
        if mode is None:
            params = stage.cached_params() if hasattr(stage, "cached_params") else ()
            held = [p for p in params if p.requires_grad and p.grad is not None]
            mode = arena._dp_deferred = bool(held)
            if mode:  # autograd will ADD this arena into the existing p.grad: reduce p.grad itself, afterwards
                self._deferred.extend(p for p in params if p.requires_grad)
        if not mode:
            start = getattr(arena, "_dp_sent", 0)  # progress lives on the arena itself
            final = upto >= arena.total
            if upto - start <= 0 or (upto - start < self.min_bucket and not final):
                return
            arena._dp_sent = upto
            self._launch(arena.flat[start:upto], keep=arena)
        if not self._cb_queued:
            try:  # end-of-backward hook (the mechanism DDP uses); outside backward the caller runs finish()
                torch.autograd.Variable._execution_engine.queue_callback(self.finish)
                self._cb_queued = True
            except RuntimeError:
                pass

    def _launch(self, flat, keep=None):
        if dist.get_backend(self.pg) == "nccl":
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
            self._pending.append((work, flat, keep, False))
        else:  # gloo (CPU tests, and the 2-ranks-on-one-GPU parity run of tests/test_dp_gpu.py): no AVG reduction
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._pending.append((work, flat, keep, True))
        self.buckets_last_step += 1
        self.bytes_last_step += flat.numel() * flat.element_size()

    def no_sync(self):
        """Context manager: run forward/backward WITHOUT gradient communication (torch DDP's no_sync)."""
        dp = self

        class _NoSync:
            def __enter__(self):
                self.prev, dp._enabled = dp._enabled, False

            def __exit__(self, *exc):
                dp._enabled = self.prev

        return _NoSync()

    def finish(self):
        """Block the current stream until every outstanding gradient all-reduce has completed."""
        for work, flat, _keep, divide in self._pending:
            work.wait()
            if divide:
                flat.div_(self.world)
        self._pending.clear()
        self._cb_queued = False
        if self._deferred:  # runs after every AccumulateGrad of this backward (engine final callback)
            seen, grads = set(), []
            for p in self._deferred:
                if id(p) not in seen and p.grad is not None:
                    seen.add(id(p))
                    grads.append(p.grad)
            self._deferred.clear()
            if grads:
                flat = torch._utils._flatten_dense_tensors(grads)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg)
                flat.div_(self.world)
                for g, r in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                    g.copy_(r)
                self.buckets_last_step += 1
                self.bytes_last_step += flat.numel() * flat.element_size()


def shard_batch(tensors, rank, world):
    """rank r gets samples [r*B/world, (r+1)*B/world) of every tensor."""
    out = []
    for t in tensors:
        B = t.shape[0]
        if B % world:
            raise ValueError(f"global batch {B} is not divisible by world size {world}")
        per = B // world
        out.append(t[rank * per:(rank + 1) * per])
    return out


def folds_for_rank(n_folds, rank, world):
    """Cross-validation folds are independent training jobs (train_kfold.py:151-155 runs them one after another):
    fold k goes to rank k mod world, no communication between them (SURVEY.md section 8e, configs[4])."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return [k for k in range(int(n_folds)) if k % world == rank]
