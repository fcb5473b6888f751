"""Data-parallel plumbing for cnn_train_dag's multi-GPU mode (emoVoxCeleb/run_distillation.m:179-181:
'gpus', opts.gpus with parameterServer.method = 'tmove'): one process per GPU, each takes an interleaved
slice of the minibatch, gradients are SUMMED across processes before the update, and the update divides
by the global batch size.  Here the exchange is an NCCL all-reduce over NVLink (gloo in the CPU tests)
on the flat fp32 gradient buffer, split into buckets so that it can be issued while the rest of the
backward pass is still running."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_batch(batch_indices, rank, world):
    """cnn_train_dag gives worker `labindex` the samples batch(labindex:numlabs:end)."""
    return list(batch_indices[rank::world])


def bucket_bounds(segments, bucket_elems):
    """Group consecutive parameter segments [(name, offset, numel)] (in buffer order) into buckets of at
    least `bucket_elems` elements; returns [(start, end)] element ranges covering the buffer exactly."""
    bounds, start, end = [], None, None
    for _, off, n in segments:
        if start is None:
            start = off
        end = off + n
        if end - start >= bucket_elems:
            bounds.append((start, end))
            start = None
    if start is not None:
        bounds.append((start, end))
    return bounds


class GradientAllReducer:
    """Sums a flat gradient tensor across the process group, bucket by bucket."""

    def __init__(self, bounds=None, group=None, async_op=False):
        self.bounds, self.group, self.async_op = bounds, group, async_op

    def __call__(self, flat):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return flat
        bounds = self.bounds or [(0, flat.numel())]
        handles = [dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=self.async_op) for a, b in bounds]
        if self.async_op:
            for h in handles:
                h.wait()
        return flat
