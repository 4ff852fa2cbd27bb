#!/usr/bin/env python
"""NCCL check of the chunk-range sharded demix (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_shard_check.py

Every rank demixes its chunk range of one synthetic track with a small fp32 BS-RoFormer, the only exchange is the
overlap-add halo (isend/irecv of raw partial sums over NVLink); rank 0 gathers the owned spans and compares them with its
own single-GPU demix of the whole track.  One chunk per network call on both sides, so the comparison is bitwise.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from audiolab_b200.configs import RoformerConfig
    from audiolab_b200.demix import RoformerDemixer
    from audiolab_b200.nets.roformer import RoformerMaskNet
    from audiolab_b200.sharding import ShardedRoformerDemixer

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(4321)
    cfg = RoformerConfig(dim=64, depth=1, heads=2, dim_head=32, chunk_size=441 * 200)     # 2 s chunks, step 0.5 s
    net = RoformerMaskNet(cfg).to(dev).eval()
    demixer = RoformerDemixer(cfg, net, batch_size=1)
    g = torch.Generator().manual_seed(99)
    n = 44100 * 21 + 1234
    mix = (torch.rand((2, n), generator=g) - 0.5).to(dev)
    span, cr = ShardedRoformerDemixer(demixer, rank, world).demix_span(mix)
    rows = cfg.num_stems * cfg.audio_channels
    full = torch.zeros((rows, n), dtype=torch.float32, device=dev)
    if span is not None:
        full[:, cr.p0:cr.p1] = span
    dist.all_reduce(full)                       # spans are disjoint: the sum is the stitched track
    torch.cuda.synchronize()
    if rank == 0:
        ref = demixer.demix(mix)
        ref = ref.reshape(rows, n) if ref.dim() == 3 else ref
        err = float((full - ref).abs().max())
        print(json.dumps({"check": "nccl chunk-range sharding vs single GPU", "world_size": world, "n_samples": n,
                          "bitwise_equal": bool(torch.equal(full, ref)), "max_abs_diff": err,
                          "ranges": "see sharding.plan_chunk_ranges", "halo_in": cr.halo_in, "halo_out": cr.halo_out}))
        assert err <= 1e-6, err
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
