#!/bin/bash
# full GPU test suite, smoke, ncu capture of the GEMM kernels, the default bench line (all configs + cpu baseline)
TAG=${1:-ev2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -4 $OUT/smoke.log
echo "== big-K gemm"; timeout 120 python - > $OUT/gemm_bigk.log 2>&1 <<'PY'
import json, torch, sys
sys.path.insert(0, '.')
from audiolab_b200 import netops
def timed(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for (m,n,k) in [(8192,8192,8192),(16384,4096,4096),(131072,2048,2048)]:
    a=torch.randn(m,k,device='cuda').bfloat16(); w=torch.randn(n,k,device='cuda').bfloat16(); o=torch.empty(m,n,device='cuda',dtype=torch.bfloat16)
    t=timed(lambda: netops.gemm_bf16(a,w,o)); t2=timed(lambda: torch.nn.functional.linear(a,w))
    print(json.dumps({"m":m,"n":n,"k":k,"ms":round(t,3),"tflops":round(2*m*n*k/t/1e9,1),"torch_ms":round(t2,3),"torch_tflops":round(2*m*n*k/t2/1e9,1)}))
PY
cat $OUT/gemm_bigk.log | tail -4
echo "== ncu gemm"; timeout 420 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -c 6 -o $OUT/ncu_gemm python tools/gpu/gemm_debug.py > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_gemm.log; ls -la $OUT/*.ncu-rep 2>/dev/null
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "rc=$?"; cut -c1-300 $OUT/bench_default.json; tail -3 $OUT/bench_default.err
