"""Multi-GPU partitioning of the demix path.  One process per GPU.

* batch of tracks (BASELINE cfg 4): independent units -> ``assign_tracks`` (longest-first greedy);
  no data-path collective.
* one long track (cfg 5): rank r owns a contiguous chunk range and the matching output span.  Chunks
  are independent until the overlap-add; the last chunks of rank r spill ``chunk - step`` samples into
  rank r+1's span, so each interior boundary needs ONE neighbour exchange of the left rank's raw
  partial sums (rows x (chunk-step) floats ~ 2.1 MB) -- ``torch.distributed`` send/recv (NCCL over
  NVLink on the GPU box, gloo in the CPU tests).  Weight sums are analytic and recomputed locally.
  The receiver continues the left-to-right sum from the received partial sums, so the sharded
  result is bit-identical to the single-GPU result.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def assign_tracks(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-first greedy: returns, per rank, the indices of the tracks it processes."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += lengths[i]
    for r in range(world_size):
        out[r].sort()
    return out


@dataclass
class ChunkRange:
    c0: int          # first owned chunk
    c1: int          # one past the last owned chunk
    p0: int          # owned output span [p0, p1)
    p1: int
    halo_in: int     # samples at the start of the span that receive the left neighbour's partial sums
    halo_out: int    # samples past p1 this rank's chunks reach (sent to the right neighbour)


def plan_chunk_ranges(offsets: Sequence[int], chunk_len: int, n_total: int, world_size: int) -> List[ChunkRange]:
    """Contiguous, balanced chunk ranges, one per rank.

    Every range except the last live one is extended until the next range starts at least `chunk_len`
    past its own first offset, so a rank's chunks never reach beyond its right neighbour's span and a
    single neighbour exchange suffices.  Ranks left without chunks get empty ranges (c0 == c1).
    """
    n = len(offsets)
    live: List[Tuple[int, int]] = []
    c = 0
    for r in range(world_size):
        if c >= n:
            break
        take = -(-(n - c) // (world_size - r))
        c1 = min(n, c + take)
        while c1 < n and offsets[c1] - offsets[c] < chunk_len:
            c1 += 1
        live.append((c, c1))
        c = c1
    if c < n:
        live[-1] = (live[-1][0], n)
    out: List[ChunkRange] = []
    for i, (a, b) in enumerate(live):
        last = i == len(live) - 1
        p0 = 0 if i == 0 else offsets[a]
        p1 = n_total if last else offsets[live[i + 1][0]]
        halo_in = 0 if i == 0 else max(0, min(n_total, offsets[a - 1] + chunk_len, p1) - p0)
        halo_out = 0 if last else max(0, min(n_total, offsets[b - 1] + chunk_len) - p1)
        out.append(ChunkRange(a, b, p0, p1, halo_in, halo_out))
    for i in range(len(out) - 1):
        if out[i].halo_out > out[i + 1].p1 - out[i + 1].p0:
            raise ValueError("chunk ranges too short for a single neighbour exchange; use fewer ranks")
    while len(out) < world_size:
        out.append(ChunkRange(n, n, n_total, n_total, 0, 0))
    return out


def start_halo_exchange(halo_out: Optional[torch.Tensor], halo_in: Optional[torch.Tensor], right: Optional[int],
                        left: Optional[int], group=None) -> list:
    """Post one neighbour exchange without waiting: send `halo_out` to `right`, receive `halo_in` from `left` (in
    place).  Both operations go out as ONE group (a single NCCL kernel per rank, no send-before-receive chain along
    the ranks); returns the requests to `wait()` on before `halo_in` is read."""
    ops = []
    if right is not None and halo_out is not None and halo_out.numel():
        ops.append(dist.P2POp(dist.isend, halo_out, right, group))
    if left is not None and halo_in is not None and halo_in.numel():
        ops.append(dist.P2POp(dist.irecv, halo_in, left, group))
    return list(dist.batch_isend_irecv(ops)) if ops else []


def exchange_halo(halo_out: Optional[torch.Tensor], halo_in: Optional[torch.Tensor], rank: int,
                  right: Optional[int], left: Optional[int], group=None) -> None:
    """Blocking form of `start_halo_exchange`."""
    for req in start_halo_exchange(halo_out, halo_in, right, left, group):
        req.wait()


def sharded_ola(chunk_waves_fn: Callable[[int, int], torch.Tensor], gather_fn: Callable[..., torch.Tensor],
                offsets: Sequence[int], chunk_len: int, n_total: int, rows: int, rank: int, world_size: int,
                device, group=None, stats: Optional[dict] = None) -> Tuple[Optional[torch.Tensor], ChunkRange]:
    """Chunk-range sharded demix + overlap-add, with the halo exchange overlapped with the interior chunks.

    chunk_waves_fn(c0, c1) -> [c1-c0, rows, chunk_len] chunk outputs of chunks [c0, c1).
    gather_fn(waves, c0, c1, p0, p1, halo_in, raw_out) -> [rows, p1-p0]: ascending-chunk weighted sum
        over chunks [0, c1) with data only for [c0, c1) (normalised unless raw_out).
    Order of work on a rank: (1) the TAIL chunks -- the ones that reach past the owned span -- are evaluated first and
    their raw partial sums over the neighbour's head are posted (isend) together with the receive of this rank's own
    head halo; (2) the interior chunks are evaluated while the exchange is in flight; (3) wait, then the ascending
    gather of the owned span continues the left neighbour's partial sums.  The order in which chunks are EVALUATED does
    not enter the result: the gather always sums in ascending chunk order.
    Returns (owned span [rows, p1-p0] or None for an idle rank, its ChunkRange).  `stats` (optional dict) receives
    halo_bytes_out / halo_bytes_in / tail_chunks.
    """
    plan = plan_chunk_ranges(offsets, chunk_len, n_total, world_size)
    me = plan[rank]
    live = [i for i, cr in enumerate(plan) if cr.c1 > cr.c0]
    if me.c1 == me.c0:
        return None, me
    pos = live.index(rank)
    left = live[pos - 1] if pos > 0 else None
    right = live[pos + 1] if pos + 1 < len(live) else None
    # first chunk whose samples reach past the owned span
    ct = me.c1
    if right is not None and me.halo_out:
        while ct > me.c0 and offsets[ct - 1] + chunk_len > me.p1:
            ct -= 1
    tail = chunk_waves_fn(ct, me.c1) if ct < me.c1 else None
    halo_out = None
    if right is not None and me.halo_out:
        halo_out = gather_fn(tail, ct, me.c1, me.p1, me.p1 + me.halo_out, None, True).contiguous()
    halo_in = torch.empty((rows, me.halo_in), dtype=torch.float32, device=device) if me.halo_in else None
    reqs = start_halo_exchange(halo_out, halo_in, right, left, group)
    if stats is not None:
        stats["halo_bytes_out"] = 0 if halo_out is None else halo_out.numel() * 4
        stats["halo_bytes_in"] = 0 if halo_in is None else halo_in.numel() * 4
        stats["tail_chunks"] = me.c1 - ct
    if ct > me.c0:
        head = chunk_waves_fn(me.c0, ct)
        waves = head if tail is None else torch.cat((head, tail), dim=0)
    else:
        waves = tail
    timed = stats is not None and stats.get("time_wait") and torch.device(device).type == "cuda" and reqs
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for req in reqs:
        req.wait()
    if timed:
        e1.record()
        stats.setdefault("wait_events", []).append((e0, e1))
    parts = []
    if me.halo_in:
        parts.append(gather_fn(waves, me.c0, me.c1, me.p0, me.p0 + me.halo_in, halo_in, False))
    if me.p0 + me.halo_in < me.p1:
        parts.append(gather_fn(waves, me.c0, me.c1, me.p0 + me.halo_in, me.p1, None, False))
    span = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    return span, me


class ShardedRoformerDemixer:
    """Chunk-range sharding of RoformerDemixer.demix across the ranks of a process group."""

    def __init__(self, demixer, rank: int, world_size: int, group=None):
        self.demixer, self.rank, self.world_size, self.group = demixer, rank, world_size, group

    @torch.no_grad()
    def demix_span(self, mix: torch.Tensor, stats: Optional[dict] = None):
        """Every rank holds `mix` [s, n]; returns (owned span [stems*s, p1-p0] or None, ChunkRange)."""
        from . import spectral as sp
        from .demix import _dev_i32, _dev_i64, roformer_schedule
        d = self.demixer
        c = d.cfg
        C, n = c.chunk_size, mix.shape[1]
        if n < C:
            raise ValueError("track shorter than one chunk: nothing to shard")
        offs, mult = roformer_schedule(n, C, c.step)
        dev = mix.device
        rows = c.num_stems * c.audio_channels
        offs_d, mult_d, wtab = _dev_i64(offs, dev), _dev_i32(mult, dev), d.window(dev)

        def waves_fn(c0, c1):
            return d.chunk_waves(mix, offs[c0:c1])

        def gather_fn(waves, c0, c1, p0, p1, halo_in, raw_out):
            out = torch.empty((rows, p1 - p0), dtype=torch.float32, device=dev)
            # positions [p0, p1) land in a [rows, p1-p0] buffer (base pointer shifted by -p0)
            sp.ola_gather(waves, offs_d[:c1], n, mult=mult_d[:c1], wtab=wtab, p0=p0, p1=p1,
                                 halo_in=halo_in, raw_out=raw_out, eps=1e-10, data_chunk0=c0,
                                 out=_ShiftedOut(out, p0))
            return out

        return sharded_ola(waves_fn, gather_fn, offs, C, n, rows, self.rank, self.world_size, dev, self.group, stats)


class _ShiftedOut:
    """Marker understood by spectral.ola_gather: write track position p at column p - shift of `buf`."""

    def __init__(self, buf: torch.Tensor, shift: int):
        self.buf, self.shift = buf, shift
