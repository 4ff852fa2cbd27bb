#!/bin/bash
# kernel table of one bench step (torch profiler) + ncu --set full of the band-axis attention kernel at the bench shape
TAG=${1:-band}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/gpu/step_profile.py > $OUT/step_profile.txt 2>&1; echo "rc=$?"; grep -v Warning $OUT/step_profile.txt | cut -c1-170 | head -22
cat > /tmp/ba_target.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from audiolab_b200 import netops
n_seq, F, H = 27 * 801, 62, 8
q, k, v = (torch.randn(n_seq * F, H * 64, device="cuda").half() for _ in range(3))
gates = torch.randn(n_seq * F, 16, device="cuda").half()[:, :H]
for _ in range(2):
    netops.band_attention(q, k, v, n_seq, F, H, 64, gates=gates)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    netops.band_attention(q, k, v, n_seq, F, H, 64, gates=gates)
e1.record(); torch.cuda.synchronize()
print("band attention ms", e0.elapsed_time(e1) / 10)
PY
timeout 120 python /tmp/ba_target.py 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:band_attn -s 1 -c 1 -f -o /tmp/ncu_ba python /tmp/ba_target.py > $OUT/ncu_ba.log 2>&1; echo "rc=$?"
ncu -i /tmp/ncu_ba.ncu-rep --page raw --csv > $OUT/ncu_ba_raw.csv 2>/dev/null
ncu -i /tmp/ncu_ba.ncu-rep --page source --csv > $OUT/ncu_ba_source.csv 2>/dev/null
python tools/ncu_raw_extract.py $OUT/ncu_ba_raw.csv | head -24
python tools/ncu_hot_sass.py $OUT/ncu_ba_source.csv band_attn 0 25
gzip -f $OUT/ncu_ba_source.csv
