"""CPU (gloo, world_size 2): the multi-GPU host logic -- frame ranges, all-gather of fixed-stride pose tables,
rank-order merge with sample-id offsets, capacity-flag propagation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_table(rank, n, cap, flags=0):
    from fastposecnn_b200 import _lib
    t = torch.zeros((cap + 1, _lib.POSE_ROW), dtype=torch.float32)
    hdr = t[0, :_lib.NUM_COUNTERS].view(torch.int32)
    hdr[_lib.CNT_INSTANCES] = n
    hdr[_lib.CNT_FLAGS] = flags
    rows = t[1:1 + n]
    ri = rows.view(torch.int32)
    for i in range(n):
        ri[i, _lib.ROW_CLASS] = 1 + (i + rank) % 6
        ri[i, _lib.ROW_SAMPLE] = i // 2            # local frame index
        ri[i, _lib.ROW_COUNT] = 100 + i
        rows[i, _lib.ROW_XY] = 10.0 * rank + i
        rows[i, _lib.ROW_XY + 1] = 0.5
    return t


def _worker(rank, world, port, flags, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fastposecnn_b200 import _lib
    from fastposecnn_b200.sharding import gather_tables, merge_tables, shard_range
    cap = 8
    n_local = 3 if rank == 0 else 5
    local = _fake_table(rank, n_local, cap, flags if rank == 1 else 0)
    out = torch.empty((world, cap + 1, _lib.POSE_ROW), dtype=torch.float32)
    gather_tables(local, out)
    frames = [shard_range(7, r, world)[1] - shard_range(7, r, world)[0] for r in range(world)]
    try:
        agg = merge_tables(out, frames)
        res = ("ok", agg["sample_ids"].tolist(), agg["class_ids"].tolist(), agg["xy"][:, 0].tolist(), frames)
    except RuntimeError as e:
        res = ("err", str(e))
    if rank == 0:
        out_q.put(res)
    dist.destroy_process_group()


def _run(flags):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, flags, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return q.get()


def test_shard_range_covers_batch():
    from fastposecnn_b200.sharding import shard_range
    for b in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(b, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_gather_and_merge_world2():
    status, sample_ids, class_ids, xs, frames = _run(0)
    assert status == "ok" and frames == [4, 3]
    assert sample_ids == [0, 0, 1] + [4 + v for v in [0, 0, 1, 1, 2]]      # rank 1's frames start at 4
    assert class_ids == [1, 2, 3] + [2, 3, 4, 5, 6]
    assert xs == [0.0, 1.0, 2.0, 10.0, 11.0, 12.0, 13.0, 14.0]             # rank order == reference instance order


@pytest.mark.timeout(300)
def test_capacity_flag_propagates():
    res = _run(1)
    assert res[0] == "err" and "FPC_ECAPACITY" in res[1] and "rank 1" in res[1]
