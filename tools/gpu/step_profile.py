"""Where one bench step goes: kernel-name table of one demix of the contract workload under torch.profiler (CUPTI
activity records -- in-process, no replays; the timed number of a profiled run is never a bench value)."""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from torch.profiler import ProfilerActivity, profile

    import bench
    from audiolab_b200.separator import Separator
    dev = torch.device("cuda:0")
    sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev),
                    mdxc_params={"batch_size": 27, "overlap": 4})
    inst = sep.load_model(sys.argv[1] if len(sys.argv) > 1 else "model_bs_roformer_ep_368_sdr_12.9628.ckpt")
    mix = torch.from_numpy(bench.synth_mix(60 * bench.SR, seed=1236)).to(dev)
    for _ in range(2):
        inst.demixer.demix(mix)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        inst.demixer.demix(mix)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            t = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            agg[ev.name][0] += 1
            agg[ev.name][1] += t
            total += t
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    print(json.dumps({"total_kernel_ms": round(total / 1e3, 2), "kernels": len(rows)}))
    for name, (n, t) in rows[:40]:
        print(f"{t / 1e3:9.3f} ms  {100 * t / total:5.1f} %  x{n:<5d} {name[:150]}")


if __name__ == "__main__":
    main()
