"""SI-SDR of the bf16 mask-network paths at the REAL size (BS-RoFormer dim 512, depth 12, 62 bands; one 8 s chunk,
weights seed 4321, mix seed 1236) against the fp32 CPU oracle -- the table VERDICT r1 asked for.  Writes JSON lines.

  tc       : tcgen05 GEMM path (bf16 MMA operands, fp32 residual stream + fp32 norm statistics)
  tc_fp16  : the same path with IEEE-half operands (11 significand bits instead of 8, same tensor-core rate)
  fused    : round-1 path (cuBLAS, bf16 residual stream)                     AUDIOLAB_B200_TC_GEMM=0 semantics
  fp32     : the CUDA fp32 path (parity configuration)
  oracle16 : the ORACLE itself under bf16 autocast on the CPU (the reference's use_autocast=True arithmetic)
"""
import dataclasses
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import roformer as oro  # noqa: E402
from oracle.metrics import max_abs_err, si_sdr_db  # noqa: E402
from oracle.synth import synth_mix  # noqa: E402


def main():
    import audiolab_b200.nets.roformer as rof
    from audiolab_b200.configs import RoformerConfig
    from audiolab_b200.demix import RoformerDemixer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_num_threads(os.cpu_count() or 1)
    depth = int(os.environ.get("PARITY_DEPTH", "12"))
    oc = oro.RoformerConfig(depth=depth)
    om = oro.build_roformer(oc, seed=4321)
    mix = torch.tensor(synth_mix(oc.chunk_size, seed=1236))
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = oro.demix_roformer(mix, om, oc)
    print(json.dumps({"oracle_fp32_s": round(time.perf_counter() - t0, 2), "depth": depth}), flush=True)
    pc = RoformerConfig(**dataclasses.asdict(oc))
    out = {}
    paths = (("tc", torch.bfloat16, True), ("tc_fp16", torch.float16, True), ("fused", torch.bfloat16, False),
             ("fp32", torch.float32, True))
    for name, dtype, tc in paths:
        pm = rof.RoformerMaskNet(pc)
        pm.load_state_dict(om.state_dict(), strict=True)
        pm = pm.cuda().eval().set_compute_dtype(dtype)
        rof._TC_GEMM = tc
        d = RoformerDemixer(pc, pm, batch_size=1)
        got = d.demix(mix.cuda()).cpu()
        out[name] = got
        print(json.dumps({"path": name, "si_sdr_db_vs_oracle_fp32": round(si_sdr_db(got, ref), 2),
                          "max_abs_err": float(max_abs_err(got, ref)), "ref_peak": float(ref.abs().max())}), flush=True)
        del pm, d
        torch.cuda.empty_cache()
    rof._TC_GEMM = True
    if "--no-oracle16" not in sys.argv:
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            r16 = oro.demix_roformer(mix, om, oc).float()
        print(json.dumps({"path": "oracle16", "si_sdr_db_vs_oracle_fp32": round(si_sdr_db(r16, ref), 2),
                          "seconds": round(time.perf_counter() - t0, 1),
                          "tc_vs_oracle16_db": round(si_sdr_db(out["tc"], r16), 2)}), flush=True)


if __name__ == "__main__":
    main()
