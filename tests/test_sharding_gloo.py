"""Multi-process (world_size 2, gloo, CPU) test of the N>1 host logic: frame-range sharding,
budget closed form, result gather and output order. The per-rank "encoder" here is the CPU
oracle (tests may use it); on GPUs the same code runs with NCCL and the CUDA encoder
(tests/test_gpu_multi.py, bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

from psxavenc_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_ranges_cover_in_order():
    for n in (0, 1, 7, 4096, 10000):
        for world in (1, 2, 3, 8):
            ranges = [sharding.frame_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_budget_closed_form_matches_accumulator():
    """mdec.c:772-774 run sequentially vs the per-rank closed form (strcd: 1050/120 sectors)."""
    for num, den in ((150, 15), (1050, 120), (131, 30), (75, 24)):
        acc, seq = 0, []
        for _ in range(500):
            acc += num
            seq.append(acc // den * 2016)
            acc %= den
        assert list(sharding.frame_budgets(500, num, den)) == seq
        assert list(sharding.frame_budgets(100, num, den, first_frame=400)) == seq[400:]
    assert list(sharding.frame_budgets(4, 1050, 120)) == [16128, 18144, 18144, 18144]


def test_stream_ownership():
    for world in (1, 2, 8):
        owned = [sharding.streams_of(r, 19, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(19))
        assert all(sharding.stream_owner(s, world) == r for r in range(world) for s in owned[r])


def _worker(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h, n = 64, 48, 11
        frames = np.stack([synth.gen_frame(i, w, h, 3 + i % 3) for i in range(n)])
        budgets = sharding.frame_budgets(n, 1050, 480)          # mixes 4032- and 6048-byte frames
        first, last = sharding.frame_range(n, rank, world)
        orc = oracle.Restated()
        out, res = orc.bs_encode_batch(0, w, h, frames[first:last], budgets[first:last], oracle.FDCT_ISLOW, stride=6048)
        all_res = sharding.gather_results(torch.from_numpy(res), n, dist)
        outs = [None] * world
        dist.all_gather_object(outs, out)
        if rank == 0:
            full_out, full_res = orc.bs_encode_batch(0, w, h, frames, budgets, oracle.FDCT_ISLOW, stride=6048)
            ok = bool(np.array_equal(all_res.numpy(), full_res) and np.array_equal(np.concatenate(outs), full_out))
            queue.put(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_encode_matches_single():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert queue.get(timeout=5) is True
