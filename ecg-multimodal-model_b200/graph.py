"""One training step as ONE CUDA graph.

The reference loop (train.py:60-86) launches its step kernel by kernel from Python; so does this package in eager
mode -- about 400 C-ABI calls, ~10 ms of host time per step, which is more than half of the GPU time of a step at the
per-GPU batch of the 8-GPU configuration and sits in front of every step once the loop reads `loss.item()`.
GraphedTrainStep captures zero_grad + forward + loss + backward (+ the bucketed gradient all-reduce of
ecgmm.parallel.DataParallel) + optimizer.step() for a fixed batch shape and replays it with a single launch:

    step = ecgmm.graph.GraphedTrainStep(model_or_dp, criterion, optimizer, example_batch)      # once
    for images, ecg, clin, labels in loader:
        loss = step(images, ecg, clin, labels)         # copies the batch into the graph's input buffers, replays
        running_loss += loss.item()

What has to differ from a naive capture (kernel arguments are frozen at capture time):
  * Adam's step count and every param group's learning rate are read from device memory (ecgmm_adam_step_dev); the
    count is advanced by a kernel at the head of the graph, `param_group['lr']` changes (train.py:158-161, OneCycleLR)
    are copied to the device before the next replay;
  * dropout seeds get a per-replay offset from the same device word (ecgmm_dropout_fwd seed_dev);
  * BatchNorm running statistics / num_batches_tracked were device-side already.
The wrapper keeps `optimizer.state[p]['step']` and the parameters' version counters in step with the replays, so
state_dict(), checkpoints and eager evaluation passes in between behave as in the eager loop.
The batch shape (and the module's train()/requires_grad configuration) is fixed per instance; build another instance
for another shape.  Memory: the graph owns the activations of one step for its lifetime.
"""
from __future__ import annotations

import contextlib
import gc

import torch

from . import lib, ops


@contextlib.contextmanager
def _quiet_gc():
    """No garbage collection while a stream capture is running: a collected object may own pinned host memory, and
    torch's pinned allocator records an event on every stream that used a block when the block is freed -- on a
    capturing stream that is a captured event which cannot be queried later (cudaErrorInvalidValue at some unrelated
    Tensor.item())."""
    was = gc.isenabled()
    gc.collect()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def _default_loss(criterion):
    # train.py:72,78: CrossEntropy(fusion_logits) + 0.1 * var_loss on the 6-tuple; plain criterion on a single tensor
    def fn(out, labels):
        if isinstance(out, (tuple, list)):
            return criterion(out[3], labels) + 0.1 * out[4]
        return criterion(out, labels)

    return fn


class GraphedTrainStep:
    def __init__(self, net, criterion, optimizer, example_batch, loss_fn=None, warmup_steps=1, restore=True):
        if not isinstance(optimizer, torch.optim.Optimizer) or not hasattr(optimizer, "_graph_mode"):
            raise lib.EcgmmError("GraphedTrainStep needs an ecgmm.optim.Adam optimizer")
        batch = list(example_batch)
        if len(batch) < 2 or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in batch):
            raise lib.EcgmmError("example_batch must be CUDA tensors (inputs..., labels)")
        self.net, self.optimizer = net, optimizer
        self.loss_fn = loss_fn if loss_fn is not None else _default_loss(criterion)
        self.static = [t.detach().clone() for t in batch]
        self.params = [p for g in optimizer.param_groups for p in g["params"]]
        dev = batch[0].device

        def eager_step():
            optimizer.zero_grad(set_to_none=True)
            out = net(*self.static[:-1])
            loss = self.loss_fn(out, self.static[-1])
            loss.backward()
            optimizer.step()
            return loss

        # The warm-up below really trains on the example batch; with restore=True parameters, buffers and optimizer
        # state are put back afterwards (in place: the graph has captured their addresses).
        module = net.module if hasattr(net, "module") and isinstance(getattr(net, "module"), torch.nn.Module) else net
        snap = None
        if restore:
            tensors = list(module.parameters()) + list(module.buffers())
            snap = ([t.detach().clone() for t in tensors], tensors,
                    {p: {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
                     for p, st in optimizer.state.items()})

        # eager warm-up on a side stream (allocator, Adam state, weight shadows, NCCL communicators)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, int(warmup_steps))):
                eager_step()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        optimizer.zero_grad(set_to_none=True)
        if snap is not None:
            with torch.no_grad():
                for saved, t in zip(snap[0], snap[1]):
                    t.copy_(saved)
                for p in self.params:
                    st = optimizer.state.get(p)
                    if not st:
                        continue
                    old = snap[2].get(p)
                    for k, v in st.items():
                        if isinstance(v, torch.Tensor):
                            v.copy_(old[k]) if old else v.zero_()
                        else:
                            st[k] = old[k] if old else 0
            torch.autograd.graph.increment_version(self.params)
        optimizer.reserve_tables()
        torch.cuda.empty_cache()

        steps = {int(optimizer.state[p]["step"]) for p in self.params if p in optimizer.state and len(optimizer.state[p])}
        if len(steps) != 1:
            raise lib.EcgmmError(f"all parameters must have taken the same number of optimizer steps, got {sorted(steps)}")
        # device state: [step count, dropout seed offset]; one fp32 learning rate per param group
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)
        self.state[0] = steps.pop()
        self.lr_host = [float(g["lr"]) for g in optimizer.param_groups]
        self.lr_dev = torch.tensor(self.lr_host, dtype=torch.float32, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        ops.GRAPH_STATE = self.state
        optimizer._graph_mode = (self.state, self.lr_dev)
        try:
            with _quiet_gc(), torch.cuda.graph(self.graph):
                lib.call("ecgmm_step_advance", ops._ptr(self.state), ops._s())
                self.loss = eager_step()
        finally:
            ops.GRAPH_STATE = None
            optimizer._graph_mode = None
        # capture executed the Python side of one step but no kernel: take that step back on the host
        for p in self.params:
            st = optimizer.state.get(p)
            if st:
                st["step"] = int(st["step"]) - 1
        self.replays = 0

    def __call__(self, *batch):
        if len(batch) == 1 and isinstance(batch[0], (tuple, list)):
            batch = tuple(batch[0])
        if len(batch) != len(self.static):
            raise lib.EcgmmError(f"expected {len(self.static)} tensors, got {len(batch)}")
        for dst, src in zip(self.static, batch):
            if src is dst:
                continue
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise lib.EcgmmError(f"batch tensor {tuple(src.shape)} {src.dtype} does not match the captured "
                                     f"{tuple(dst.shape)} {dst.dtype}")
            dst.copy_(src, non_blocking=True)
        lrs = [float(g["lr"]) for g in self.optimizer.param_groups]
        if lrs != self.lr_host:
            self.lr_dev.copy_(torch.tensor(lrs, dtype=torch.float32).pin_memory(), non_blocking=True)
            self.lr_host = lrs
        self.graph.replay()
        self.replays += 1
        for p in self.params:
            st = self.optimizer.state.get(p)
            if st:
                st["step"] = int(st["step"]) + 1
        torch.autograd.graph.increment_version(self.params)  # eager code (weight shadows, eval passes) sees new weights
        return self.loss

    @property
    def inputs(self):
        """The graph's own input buffers: fill them directly (e.g. from a copy stream) and call step(*step.inputs)."""
        return self.static


class GraphedEvalStep:
    """The INFERENCE step of the fusion model (train.py:183-200, train_kfold.py:80-90: model.eval(), torch.no_grad(),
    outputs = model(images, ecg_signals, clinical)) as one CUDA graph for a fixed batch shape.

        model.eval()
        infer = ecgmm.graph.GraphedEvalStep(model, (images, ecg, clinical))       # once
        outputs = infer(images, ecg, clinical)        # the reference's 6-tuple (or fusion logits for fusion_only)

    Eval mode has no per-step state (frozen BatchNorm statistics, no dropout), so nothing has to move to device
    memory; the graph is re-captured when a parameter or buffer of the module has changed since (a training epoch in
    between).  The returned tensors are the graph's output buffers: they are overwritten by the next call."""

    def __init__(self, net, example_inputs):
        inputs = list(example_inputs)
        if not inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in inputs):
            raise lib.EcgmmError("example_inputs must be CUDA tensors")
        self.net = net
        self.static = [t.detach().clone() for t in inputs]
        self.graph, self.outputs, self._versions = None, None, None

    def _module(self):
        return self.net.module if hasattr(self.net, "module") and isinstance(self.net.module, torch.nn.Module) else self.net

    def _state(self):
        m = self._module()
        return tuple((t.data_ptr(), t._version) for t in list(m.parameters()) + list(m.buffers()))

    def _capture(self):
        if self._module().training:
            raise lib.EcgmmError("GraphedEvalStep captures eval-mode statistics: call model.eval() first")
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():  # warm-up: allocator, weight shadows, kernel attributes
            self.net(*self.static)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with _quiet_gc(), torch.no_grad(), torch.cuda.graph(self.graph):
            self.outputs = self.net(*self.static)
        self._versions = self._state()

    def __call__(self, *inputs):
        if len(inputs) == 1 and isinstance(inputs[0], (tuple, list)):
            inputs = tuple(inputs[0])
        if len(inputs) != len(self.static):
            raise lib.EcgmmError(f"expected {len(self.static)} tensors, got {len(inputs)}")
        if self.graph is None or self._versions != self._state():
            self._capture()
        for dst, src in zip(self.static, inputs):
            if src is dst:
                continue
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise lib.EcgmmError(f"input {tuple(src.shape)} {src.dtype} does not match the captured "
                                     f"{tuple(dst.shape)} {dst.dtype}")
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs

    @property
    def inputs(self):
        """The graph's own input buffers: fill them directly and call infer(*infer.inputs)."""
        return self.static
