"""N>1 host logic on CPU: world_size-2 gloo run of the chunk-range sharding + halo exchange, with a
numpy stand-in for al_ola_gather that follows the same left-to-right summation contract.  The sharded
result must be BIT-identical to the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiolab_b200.demix import hamming_sym, roformer_schedule
from audiolab_b200.sharding import assign_tracks, plan_chunk_ranges, sharded_ola


def np_gather(chunks, data_chunk0, offsets, mult, w, n_total, p0, p1, halo_in, raw_out, eps=1e-10):
    """Reference semantics of al_ola_gather (include/audiolab_b200.h), float32, ascending chunks."""
    rows, C = chunks.shape[1], chunks.shape[2]
    out = np.zeros((rows, p1 - p0), np.float32)
    for p in range(p0, p1):
        acc = halo_in[:, p - p0].copy() if halo_in is not None else np.zeros(rows, np.float32)
        wsum = np.float32(0)
        for c, off in enumerate(offsets):
            if off > p:
                break
            j = p - off
            if j >= min(C, n_total - off):
                continue
            for _ in range(mult[c]):
                if c >= data_chunk0:
                    acc = (acc + chunks[c - data_chunk0, :, j] * w[j]).astype(np.float32)
                wsum = np.float32(wsum + w[j])
        out[:, p - p0] = acc if raw_out else acc / max(wsum, np.float32(eps))
    return out


def _problem():
    C, step, n, rows = 40, 10, 333, 2
    offs, mult = roformer_schedule(n, C, step)
    rs = np.random.RandomState(0)
    chunks = rs.standard_normal((len(offs), rows, C)).astype(np.float32)
    w = hamming_sym(C)
    return C, step, n, rows, offs, mult, chunks, w


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        C, step, n, rows, offs, mult, chunks, w = _problem()

        def waves_fn(c0, c1):
            return torch.from_numpy(chunks[c0:c1])

        def gather_fn(waves, c0, c1, p0, p1, halo_in, raw_out):
            h = None if halo_in is None else halo_in.numpy()
            return torch.from_numpy(np_gather(waves.numpy(), c0, offs[:c1], mult[:c1], w, n, p0, p1, h, raw_out))

        span, cr = sharded_ola(waves_fn, gather_fn, offs, C, n, rows, rank, world, torch.device("cpu"))
        ret[rank] = (None if span is None else span.numpy(), cr.p0, cr.p1)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_chunk_range_sharding_bitwise_equals_single(world):
    C, step, n, rows, offs, mult, chunks, w = _problem()
    single = np_gather(chunks, 0, offs, mult, w, n, 0, n, None, False)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    stitched = np.zeros_like(single)
    covered = 0
    for r in range(world):
        span, p0, p1 = ret[r]
        if span is not None:
            stitched[:, p0:p1] = span
            covered += p1 - p0
    assert covered == n
    assert np.array_equal(stitched, single)


def test_plan_chunk_ranges_properties():
    for n_total, C, step, world in [(158760000, 352800, 88200, 8), (3000000, 352800, 88200, 4),
                                    (400000, 352800, 88200, 8), (352800, 352800, 88200, 2)]:
        offs, _ = roformer_schedule(n_total, C, step)
        plan = plan_chunk_ranges(offs, C, n_total, world)
        assert len(plan) == world
        live = [p for p in plan if p.c1 > p.c0]
        assert live[0].c0 == 0 and live[-1].c1 == len(offs)
        assert live[0].p0 == 0 and live[-1].p1 == n_total
        for a, b in zip(live[:-1], live[1:]):
            assert a.c1 == b.c0 and a.p1 == b.p0
            assert a.halo_out == b.halo_in <= b.p1 - b.p0
        if n_total == 158760000:
            assert all(p.c1 - p.c0 in (225, 224, 226) for p in plan)       # 1800 chunks over 8 ranks
            assert all(p.halo_in == C - step for p in plan[1:])            # 264 600 samples = 2.1 MB/stem


def test_assign_tracks_balances_longest_first():
    lengths = [10, 9, 8, 7, 6, 5, 4, 3, 2, 1]
    parts = assign_tracks(lengths, 3)
    assert sorted(i for p in parts for i in p) == list(range(10))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 2
    assert assign_tracks([5] * 512, 8) == [list(range(r, 512, 8)) for r in range(8)]
