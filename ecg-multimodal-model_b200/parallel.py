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
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class DataParallel(torch.nn.Module):
    def __init__(self, module, process_group=None, broadcast=True, min_bucket_elems=0):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.min_bucket = int(min_bucket_elems)
        self._pending = []
        self._cb_queued = False
        self._enabled = True
        self.buckets_last_step = 0
        self.bytes_last_step = 0
        if broadcast and self.world > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0, group=process_group)
        for st in module.stages():
            object.__setattr__(st, "_grad_ready_cb", self._on_ready)

    def forward(self, *a, **k):
        self._pending.clear()
        self._cb_queued = False
        self.buckets_last_step = 0
        self.bytes_last_step = 0
        return self.module(*a, **k)

    # ---- called by the stages during backward
    def _on_ready(self, arena, upto):
        if self.world <= 1 or not self._enabled:
            return
        start = getattr(arena, "_dp_sent", 0)  # progress lives on the arena itself (arenas are per-backward objects)
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
        if flat.is_cuda:
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
            self._pending.append((work, flat, keep, False))
        else:  # gloo (CPU tests): no AVG reduction
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
