#!/bin/bash
# ncu --set full + source page of the time-axis attention kernel at the bench shape
TAG=${1:-fattn_ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cat > /tmp/fa_target.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from audiolab_b200 import netops
B, T, I, H = 27, 801, 62, 8
q, k, v = (torch.randn(B * T * I, H * 64, device="cuda").half() for _ in range(3))
gates = torch.randn(B * T * I, 16, device="cuda").half()[:, :H]
for _ in range(2):
    netops.time_attention(q, k, v, B, T, I, H, 64, gates=gates)
torch.cuda.synchronize()
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:time_attn_kernel -s 1 -c 1 -f -o /tmp/ncu_fa python /tmp/fa_target.py > $OUT/ncu.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu.log
ncu -i /tmp/ncu_fa.ncu-rep --page raw --csv > $OUT/ncu_fa_raw.csv 2>/dev/null
ncu -i /tmp/ncu_fa.ncu-rep --page source --csv > $OUT/ncu_fa_source.csv 2>/dev/null
python tools/ncu_raw_extract.py $OUT/ncu_fa_raw.csv > $OUT/ncu_fa_summary.txt 2>&1; cat $OUT/ncu_fa_summary.txt | head -30
python tools/ncu_hot_sass.py $OUT/ncu_fa_source.csv time_attn 0 40 > $OUT/ncu_fa_hot.txt 2>&1; head -45 $OUT/ncu_fa_hot.txt
gzip -f $OUT/ncu_fa_source.csv
