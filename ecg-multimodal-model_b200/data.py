"""Input pipeline glue for the training / evaluation loops.

The reference loop moves every batch with a blocking `.to(Config.device)` of fp32 tensors right before the forward
pass (train.py:60-63): 7.5 MB per native image, the GPU idle while the copy runs.  `Prefetcher` wraps the same loader
and hands the loop DEVICE batches whose host->device copies ran on a copy stream underneath the previous step:

    loader = DataLoader(dataset, batch_size=..., pin_memory=True)          # as in train.py, + pin_memory
    for images, ecg, clinical, labels in ecgmm.data.Prefetcher(loader, device):
        outputs = model(images, ecg, clinical)                              # unchanged loop body
        ...

* pinned staging: tensors that are not pinned yet are staged through reusable pinned buffers (one set per slot);
* `depth` device buffer sets (default 2): the copy of batch i+1 overlaps the compute of batch i; a slot is reused only
  after the consumer's stream has passed the point where it asked for the next batch (CUDA events, no host sync);
* `images_as_uint8=True`: float images in [-1, 1] produced by ToTensor + Normalize(0.5, 0.5) (dataset.py:119-123) are
  the 8-bit pixel values in disguise; they are converted back to uint8 on the host (exactly, checked) so that a quarter
  of the bytes crosses PCIe, and the model's first kernel applies the normalisation again (bit-identical to the host
  transform, tests/test_fusion_gpu.py::test_uint8_images_equal_normalised_tensors).  Loaders that already yield uint8
  images (skip the transform in the Dataset) need no flag.

No torch operator touches the data on the device; the copies are cudaMemcpyAsync on a dedicated stream.
"""
from __future__ import annotations

import torch

from . import lib


def images_to_uint8(image: torch.Tensor, check: bool = True) -> torch.Tensor:
    """Inverse of ToTensor + Normalize(0.5, 0.5): float [-1, 1] -> uint8, exact for tensors that came from 8-bit
    pixels (raises otherwise when check=True)."""
    if image.dtype == torch.uint8:
        return image
    q = (image.float() * 127.5 + 127.5).round_().clamp_(0, 255).to(torch.uint8)
    if check:
        back = (q.float() / 255.0 - 0.5) / 0.5
        if not torch.equal(back, image.float()):
            raise lib.EcgmmError("images_as_uint8: the image tensor is not ToTensor+Normalize(0.5,0.5) of 8-bit pixels; "
                                 "feed it as it is (fp32 / bf16) instead")
    return q


class Prefetcher:
    """Iterates over `loader` (any iterable of tuples / lists of CPU tensors; non-tensor items are passed through)
    and yields the same tuples with the tensors on `device`."""

    def __init__(self, loader, device=None, depth: int = 2, images_as_uint8: bool = False, image_index: int = 0,
                 check_uint8: bool = True):
        if depth < 1:
            raise lib.EcgmmError("Prefetcher depth must be >= 1")
        self.loader = loader
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise lib.EcgmmError("Prefetcher needs a CUDA device (no CPU fallback)")
        self.depth = int(depth)
        self.as_u8, self.image_index, self.check_u8 = bool(images_as_uint8), int(image_index), bool(check_uint8)
        self._copy_stream = None
        self._slots = None
        self.h2d_bytes = 0      # bytes copied so far (bench.py reports them per step)
        self.batches = 0

    def __len__(self):
        return len(self.loader)

    # ---- slot bookkeeping
    def _ensure(self):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._slots = [{"dev": None, "pin": None, "ready": torch.cuda.Event(), "free": None}
                           for _ in range(self.depth)]

    def _stage(self, slot, batch):
        """Enqueue the H2D copies of one host batch into slot's device buffers (copy stream)."""
        items = list(batch) if isinstance(batch, (tuple, list)) else [batch]
        if self.as_u8 and isinstance(items[self.image_index], torch.Tensor):
            items[self.image_index] = images_to_uint8(items[self.image_index], self.check_u8)
        s = self._slots[slot]
        sig = tuple((tuple(t.shape), t.dtype) if isinstance(t, torch.Tensor) else None for t in items)
        if s["dev"] is None or s.get("sig") != sig:  # first use, or a ragged last batch: (re)allocate this slot
            s["dev"] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) if isinstance(t, torch.Tensor) else None
                        for t in items]
            s["pin"] = [None] * len(items)
            s["sig"] = sig
        cs = self._copy_stream
        if s["free"] is not None:
            cs.wait_event(s["free"])  # the consumer has moved past the batch that lived in this slot
        out = []
        with torch.cuda.stream(cs):
            for i, t in enumerate(items):
                if not isinstance(t, torch.Tensor):
                    out.append(t)
                    continue
                if t.is_cuda:
                    src = t
                elif t.is_pinned():
                    src = t
                else:
                    if s["pin"][i] is None:
                        s["pin"][i] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    elif s["free"] is not None:
                        s["free"].synchronize()  # the staging buffer's previous copy must have left the host
                    s["pin"][i].copy_(t)
                    src = s["pin"][i]
                s["dev"][i].copy_(src, non_blocking=True)
                self.h2d_bytes += t.numel() * t.element_size()
                out.append(s["dev"][i])
            s["ready"].record(cs)
        s["host_refs"] = items  # keep pinned sources alive until the copy has run
        self.batches += 1
        return tuple(out)

    def __iter__(self):
        self._ensure()
        it = iter(self.loader)
        pending = []  # (slot, device batch) in flight, oldest first
        slot = 0
        try:
            for _ in range(self.depth):
                pending.append((slot, self._stage(slot, next(it))))
                slot = (slot + 1) % self.depth
        except StopIteration:
            it = None
        while pending:
            cur_slot, dev_batch = pending.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._slots[cur_slot]["ready"])
            yield dev_batch
            # the consumer is back: everything it enqueued on its stream so far used dev_batch
            ev = self._slots[cur_slot]["free"] or torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._slots[cur_slot]["free"] = ev
            if it is not None:
                try:
                    pending.append((cur_slot, self._stage(cur_slot, next(it))))
                except StopIteration:
                    it = None
