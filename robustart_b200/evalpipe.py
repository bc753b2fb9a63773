"""Corrupt -> classify -> count: the evaluation loop body of the hot path as one device pipeline.

Replaces, per batch, AddNoise('imagenet-c') (add_noise_utils.py:22-31) + ToTensor/Normalize
(imagenet_dataloader.py:78-79) + model forward + softmax/topk/dump (cls_solver.py:404-428) +
ImageNetEvaluator.eval (imagenet_evaluator.py:49-67) by: 1 corruption kernel -> the captured forward graph
(normalisation fused into the stem gather) -> 1 counter kernel.  The only cross-rank traffic of a whole
evaluation is ONE all-reduce of the int64 counters [top1, top5, count], instead of the
reference's per-image JSON files + merge.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import nets, ops


class CorruptEvalPipeline:
    def __init__(self, model: nets.ResNet, batch: int, h: int = 224, w: int = 224, seed: int = 0):
        self.model, self.batch, self.seed = model, batch, seed
        dev = model.device
        self.device = dev
        self.static_in = torch.zeros((batch, h, w, 3), dtype=torch.uint8, device=dev)
        self.run = model.graphed(self.static_in)
        self.static_in = self.run.static_in
        self.counters = torch.zeros(3, dtype=torch.int64, device=dev)
        self.pred = torch.empty(batch, dtype=torch.int64, device=dev)
        self.h_counters = torch.zeros(3, dtype=torch.int64).pin_memory()
        self.d_images = torch.empty_like(self.static_in)
        self.d_labels = torch.empty(batch, dtype=torch.int64, device=dev)
        self.images_done = 0
        self.launches_per_step = 1 + model.launches_per_forward() + 1

    def reset(self):
        self.counters.zero_()
        self.images_done = 0

    def step_device(self, images: torch.Tensor, labels: torch.Tensor, corruption, severity: int,
                    image_offset: Optional[int] = None) -> torch.Tensor:
        """images: uint8 NHWC CUDA [batch,h,w,3] (not modified); labels int64 CUDA.  Returns the logits
        (static buffer, valid until the next step).  Counters accumulate on the device."""
        if images.shape[0] != self.batch or labels.numel() != self.batch:
            raise ValueError("CorruptEvalPipeline was captured for batch %d: got %d images / %d labels (pad or build a second "
                             "pipeline for the ragged last batch)" % (self.batch, images.shape[0], labels.numel()))
        off = self.images_done if image_offset is None else image_offset
        if corruption is None:
            self.static_in.copy_(images)
        else:
            ops.corrupt_u8(images, corruption, severity, seed=self.seed, image_offset=off, out=self.static_in)
        logits = self.run(self.static_in, copy_in=False)
        ops.topk_count_(self.counters, logits, labels, self.pred)
        self.images_done += images.shape[0]
        return logits

    def step_host(self, images_pinned: torch.Tensor, labels_pinned: torch.Tensor, corruption, severity: int):
        """Host buffers in, metric out: H2D of the uint8 batch + labels, the device step, D2H of the
        running counters.  Returns the pinned int64[3] counters after a stream sync."""
        self.d_images.copy_(images_pinned, non_blocking=True)
        self.d_labels.copy_(labels_pinned, non_blocking=True)
        self.step_device(self.d_images, self.d_labels, corruption, severity)
        self.h_counters.copy_(self.counters, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_counters

    # -- double-buffered host interface: the H2D copy of step i+1 runs on a copy stream under step i's kernels --------
    def _init_pipelined(self):
        dev = self.device
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._d_images2 = [self.d_images, torch.empty_like(self.d_images)]
        self._d_labels2 = [self.d_labels, torch.empty_like(self.d_labels)]
        self._h_counters2 = [self.h_counters, torch.zeros(3, dtype=torch.int64).pin_memory()]
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]     # H2D of buffer k landed
        self._free = [torch.cuda.Event(), torch.cuda.Event()]      # buffer k's consumer kernels finished
        self._done = [torch.cuda.Event(), torch.cuda.Event()]      # counters of the step that used buffer k are on the host
        self._hk = 0

    def step_host_pipelined(self, images_pinned: torch.Tensor, labels_pinned: torch.Tensor, corruption, severity: int):
        """Same contract as step_host (pinned host batch in, counters on the host out, every step), software-pipelined:
        this step's H2D copy is issued on a copy stream and only the PREVIOUS step's counters are waited for, so the host
        stays one step ahead and the 38.5 MB copy hides under the previous step's kernels.  Returns the pinned counters as
        of the previous step (None on the first call); finish() returns the final ones."""
        if not hasattr(self, "_hk"):
            self._init_pipelined()
        k = self._hk & 1
        cur, cs = torch.cuda.current_stream(), self._copy_stream
        cs.wait_event(self._free[k])
        with torch.cuda.stream(cs):
            self._d_images2[k].copy_(images_pinned, non_blocking=True)
            self._d_labels2[k].copy_(labels_pinned, non_blocking=True)
            self._ready[k].record(cs)
        cur.wait_event(self._ready[k])
        self.step_device(self._d_images2[k], self._d_labels2[k], corruption, severity)
        self._free[k].record(cur)
        self._h_counters2[k].copy_(self.counters, non_blocking=True)
        self._done[k].record(cur)
        self._hk += 1
        if self._hk >= 2:
            self._done[k ^ 1].synchronize()
            return self._h_counters2[k ^ 1]
        return None

    def finish(self):
        """Drain the pipelined interface: counters after the last submitted step (pinned int64[3])."""
        if not getattr(self, "_hk", 0):
            return self.h_counters
        k = (self._hk - 1) & 1
        self._done[k].synchronize()
        return self._h_counters2[k]

    def allreduce_counters(self):
        """The single collective of an evaluation: sum the counters over ranks (NCCL over NVLink)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.counters, op=dist.ReduceOp.SUM)
        return self.counters

    @staticmethod
    def metrics(counters) -> dict:
        c = [int(v) for v in counters.tolist()]
        n = max(c[2], 1)
        return {"top1": 100.0 * c[0] / n, "top5": 100.0 * c[1] / n, "count": c[2]}
