"""Adam with the constructor / param_groups / state_dict of torch.optim.Adam (the optimizer of
train.py:43, train_kfold.py:42, signal_model.py:157), stepping every parameter of a group in ONE
multi-tensor libecgmm launch.  `param_group['lr']` may be changed between steps (train.py:158-161,
OneCycleLR in signal_model.py:158-161): lr is a kernel argument read at every step."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import lib, ops

CHUNK = 1 << 16  # elements per CTA
_CAPTURED_PINNED = []  # page-locked buffers whose copies were captured into CUDA graphs (see Adam.reserve_tables)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 grad_scale=1.0):
        if amsgrad:
            raise lib.EcgmmError("amsgrad is not implemented (the reference never enables it)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False)
        super().__init__(params, defaults)
        self.grad_scale = float(grad_scale)
        self._tables = {}
        self._graph_mode = None  # (device step state, device lr vector) while ecgmm.graph captures a step
        self._reserved = {}      # group index -> (pinned host table, device table) for captured steps
        self._retired = []       # replaced table entries kept alive across a stream capture (see _table)

    def _table(self, gi, entries):
        """Device chunk table for group gi, rebuilt only when a pointer changed."""
        key = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()) for p, g, m, v in entries)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        rows = []
        for p, g, m, v, n in key:
            for off in range(0, n, CHUNK):
                cnt = min(CHUNK, n - off)
                rows.append((p + 4 * off, g + 4 * off, m + 4 * off, v + 4 * off, cnt))
        arr = np.asarray(rows, dtype=np.int64).reshape(-1, 5)
        reserved = self._reserved.get(gi[0] if isinstance(gi, tuple) else gi)
        if self._graph_mode is not None and reserved is not None:
            # inside a stream capture nothing may be allocated on the host: use the buffers reserve_tables() pinned
            host, dev = reserved
            if len(rows) > host.shape[0]:
                raise lib.EcgmmError("reserved Adam chunk table too small")
            host[: len(rows)].copy_(torch.from_numpy(arr))
            table = dev[: len(rows)]
            table.copy_(host[: len(rows)], non_blocking=True)
        else:
            host = torch.from_numpy(arr).pin_memory()
            table = host.to(entries[0][0].device, non_blocking=True)
        old = self._tables.get(gi)
        if old is not None:
            # Do not let the replaced entry die here if a stream capture is running: its pinned host buffer belongs
            # to torch's pinned allocator, which records an event on every stream that used the block WHEN THE BLOCK IS
            # FREED.  If one of those streams is the capturing stream (torch hands out streams from a small pool, so
            # the warm-up stream of an earlier step can be the capture stream of this one) that event is a captured
            # event, and the allocator's next query of it -- at an unrelated pinned allocation such as Tensor.item()
            # -- fails with cudaErrorInvalidValue.  Retired entries are dropped at the next call outside a capture.
            self._retired.append(old)
        if self._retired and not (entries[0][0].is_cuda and entries[0][0].device.type == "cuda"
                                  and torch.cuda.is_current_stream_capturing()):
            self._retired.clear()
        self._tables[gi] = (key, table, len(rows), host)
        return table, len(rows)

    def reserve_tables(self):
        """Pin host memory and allocate device memory for every group's chunk table (called by ecgmm.graph before a
        capture, where neither allocation is allowed)."""
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                continue
            n_rows = sum((p.numel() + CHUNK - 1) // CHUNK for p in ps)
            # The H2D copy of this buffer is recorded INSIDE a stream capture.  It must not come from torch's pinned
            # allocator: that allocator tags a block with a CUDA event per asynchronous copy and queries all such
            # events at later pinned allocations (Tensor.item() makes one); an event whose last record sits in a
            # captured graph cannot be queried once that graph has been destroyed (cudaErrorInvalidValue at some
            # unrelated call, depending on when the garbage collector drops the graph).  So: page-lock our own buffer
            # (the allocator does not know it and records nothing) and never hand these few kilobytes back.
            host = torch.empty((n_rows, 5), dtype=torch.int64)
            rc = torch.cuda.cudart().cudaHostRegister(host.data_ptr(), host.numel() * 8, 0)
            if int(rc) != 0:
                raise lib.EcgmmError(f"cudaHostRegister failed with {rc}")
            dev = torch.empty((n_rows, 5), dtype=torch.int64, device=ps[0].device)
            self._reserved[gi] = (host, dev)
            _CAPTURED_PINNED.append(host)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            entries = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise lib.EcgmmError("ecgmm.optim.Adam needs CUDA parameters (no CPU fallback)")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise lib.EcgmmError("ecgmm.optim.Adam needs contiguous float32 parameters")
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.to(torch.float32).contiguous()
                    p.grad = g
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = int(st["step"]) + 1
                entries.append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            if not entries:
                continue
            steps = {int(self.state[e[0]]["step"]) for e in entries}
            b1, b2 = group["betas"]
            # parameters that joined late (different step count) get their own launch
            if self._graph_mode is not None:
                # captured step: learning rate and step count come from device memory at replay time
                if len(steps) != 1:
                    raise lib.EcgmmError("a captured step needs all parameters of a group at the same step count")
                state_dev, lr_dev = self._graph_mode
                table, n = self._table((gi, -1), entries)
                lib.call("ecgmm_adam_step_dev", ctypes.c_void_p(table.data_ptr()), n, lr_dev.data_ptr() + 4 * gi,
                         float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                         state_dev.data_ptr(), self.grad_scale, ops._s())
                torch.autograd.graph.increment_version([e[0] for e in entries])
                continue
            for ordinal, stp in enumerate(sorted(steps)):
                sub = [e for e in entries if int(self.state[e[0]]["step"]) == stp]
                # cache slot = (group, ordinal of the sub-group): the step count itself changes every step and would
                # rebuild (and leak) one table per step
                table, n = self._table((gi, ordinal if len(steps) > 1 else -1), sub)
                lib.call("ecgmm_adam_step", ctypes.c_void_p(table.data_ptr()), n, float(group["lr"]), float(b1),
                         float(b2), float(group["eps"]), float(group["weight_decay"]), stp, self.grad_scale, ops._s())
            torch.autograd.graph.increment_version([e[0] for e in entries])
        return loss
