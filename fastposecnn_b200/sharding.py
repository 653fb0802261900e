"""Multi-GPU: frames shard by image, one process per GPU, ONE collective per batch.

Frames are independent units of the path (the labelling structure never links images,
lib/aggregation_layer.py:43-59), so rank ``r`` of ``G`` runs the whole path on frames
``[r*b/G, (r+1)*b/G)``.  The only exchange is an all-gather of the fixed-stride pose table
(``[1 + max_instances, 48]`` words per rank: a header row holding the live count, then one row per
instance).  Because scipy labels image by image, concatenating the rank tables in rank order
reproduces the reference's global instance order.

The send buffer IS the engine's output buffer (``PoseRecoveryEngine.table_full``): the finalize kernel
writes pose rows, and the scan kernels write the counters, straight into it -- no pack/copy step.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .pose_recovery import table_to_agg


def shard_range(b: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous frame range of ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(b, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_tables(local_full: torch.Tensor, out: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather of ``[1+cap, 48]`` tables into ``out [world, 1+cap, 48]`` (NCCL on GPUs, gloo in CPU tests)."""
    dist.all_gather_into_tensor(out.view(-1), local_full.view(-1), group=group)
    return out


def gather_pose_tables(engine, out: torch.Tensor, group=None) -> torch.Tensor:
    return gather_tables(engine.table_full, out, group=group)


class OverlappedGather:
    """Runs the per-batch all-gather on its own stream so that the next batch's kernels (compute stream) overlap with
    it: ``after_launch(engine)`` right after ``engine.launch`` / ``replay``; ``before_reuse(engine)`` before the same
    engine is launched again; ``finish()`` before the gathered tables are read."""

    def __init__(self, engines, world: int, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.out = {id(e): torch.empty((world, e.max_instances + 1, _lib.POSE_ROW), dtype=torch.float32, device=self.device)
                    for e in engines}
        self.done = {id(e): None for e in engines}

    def after_launch(self, engine, group=None):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            gather_tables(engine.table_full, self.out[id(engine)], group=group)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.done[id(engine)] = done
        return self.out[id(engine)]

    def before_reuse(self, engine):
        done = self.done[id(engine)]
        if done is not None:
            torch.cuda.current_stream(self.device).wait_event(done)   # the table must be sent before it is overwritten

    def finish(self):
        self.stream.synchronize()


def merge_tables(gathered: torch.Tensor, frames_per_rank: List[int]) -> Dict[str, torch.Tensor]:
    """Concatenates the live rows of every rank's table in rank order and offsets ``sample_ids`` by the
    rank's first frame.  Raises if any rank reported a capacity overflow."""
    world = gathered.shape[0]
    header = gathered[:, 0, :_lib.NUM_COUNTERS].contiguous().view(torch.int32).cpu()
    parts, offset = [], 0
    for r in range(world):
        if int(header[r, _lib.CNT_FLAGS]) != 0:
            raise RuntimeError(f"libfpc_b200 error -3 (FPC_ECAPACITY): rank {r} overflowed its tables "
                               f"(flags={int(header[r, _lib.CNT_FLAGS])})")
        n = int(header[r, _lib.CNT_INSTANCES])
        rows = gathered[r, 1:1 + n]
        if n:
            rows = rows.clone()
            rows.view(torch.int32)[:, _lib.ROW_SAMPLE] += offset
        parts.append(rows)
        offset += frames_per_rank[r]
    table = torch.cat(parts, dim=0) if parts else gathered.new_zeros((0, _lib.POSE_ROW))
    return table_to_agg(table, table.shape[0])
