"""Data-parallel plumbing (one process per GPU, torch.distributed / NCCL over NVLink).

The path shards along the batch of clip windows; the only collectives are
  (1) training: one all-reduce over the flat fp32 live-gradient bucket at the end of backward -- SUM when the loss is the
      global-batch CCC (CCCLoss(global_stats=True)), MEAN for per-rank losses (see make_grad_sync),
  (2) eval / global loss: one all-reduce(sum) of the (2, 6) fp64 CCC partial sums.
`joint_modalities='NONE'` attends across the batch and is therefore per-shard
("replicas only") exactly as the reference's DataParallel scatter would make it.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> int:
    """Initialise the default process group from torchrun's env (RANK/WORLD_SIZE/MASTER_*)."""
    if dist.is_initialized():
        return dist.get_rank()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    import datetime
    # short watchdog: a mismatched collective must fail in minutes, not hang a GPU box
    dist.init_process_group(backend=backend, timeout=datetime.timedelta(seconds=int(os.environ.get("JMT_DIST_TIMEOUT_S", "180"))))
    return dist.get_rank()


def shard_bounds(n_items: int, rank: int, world: int):
    """Rank r takes windows [r*n/world, (r+1)*n/world) (contiguous, balanced to +-1)."""
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


class GradSync:
    """All-reduce of the flat fp32 gradient bucket, optionally in pieces that overlap the rest of backward.

    `start(t)` launches an asynchronous all-reduce (average) of a bucket slice whose gradients are final -- NCCL runs it
    on its own stream while the tape keeps executing (JMTPipeline starts the fusion slice before the TCN / FcLayer
    backward) -- and `finish()` makes the compute stream wait for everything started.  Calling the object with the whole
    bucket does both (the plain end-of-backward all-reduce)."""

    def __init__(self, group=None, average: bool = True):
        self.group, self.average, self.works, self.pending = group, average, [], []

    def active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def start(self, t: torch.Tensor):
        if not self.active() or t.numel() == 0:
            return
        if self.average and dist.get_backend(self.group) == "nccl":
            # NCCL averages in the collective itself (no extra pass over the bucket)
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                                   # gloo (CPU tests) has no AVG
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if self.average:
                self.pending.append(t)

    def finish(self):
        for w in self.works:
            w.wait()
        self.works = []
        for t in self.pending:
            t.mul_(1.0 / dist.get_world_size(self.group))
        self.pending = []

    def __call__(self, bucket: torch.Tensor):
        self.start(bucket)
        self.finish()


def make_grad_sync(group=None, average: Optional[bool] = None, global_loss: bool = False) -> GradSync:
    """Hook for `module.set_grad_sync`: all-reduce the flat gradient bucket in place (see GradSync).

    Which reduction is right depends on the loss:
      * per-rank loss (each rank computes the CCC of ITS shard): the data-parallel convention is the MEAN of the per-rank
        gradients -> average=True (default when global_loss is False);
      * global-batch loss (`CCCLoss(global_stats=True)`: the six sums are all-reduced inside the forward, so every rank
        back-propagates d L_global / d(its own predictions)): the gradient of that ONE loss w.r.t. the replicated
        parameters is the SUM of the per-rank pieces -> pass global_loss=True (average=False).  Averaging there would
        shrink the gradient by 1/world and make training depend on the number of GPUs."""
    if average is None:
        average = not global_loss
    return GradSync(group, average)


def allreduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce CCC partial sums ((npairs, 6) fp64) across ranks; no-op without a process group."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src=src, group=group)
