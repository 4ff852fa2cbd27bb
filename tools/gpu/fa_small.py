import sys, os, torch
sys.path.insert(0, os.getcwd())
from audiolab_b200 import netops
B, T, I, H = 1, 128, 1, 1
q, k, v = (torch.randn(B * T * I, H * 64, device="cuda").half() for _ in range(3))
o = netops.time_attention(q, k, v, B, T, I, H, 64)
torch.cuda.synchronize()
print("ok", float(o.float().abs().max()))
