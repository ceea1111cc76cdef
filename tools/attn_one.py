"""One attention shape, a few launches (ncu --set full target): python tools/attn_one.py B L heads d"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops
b, l, heads, d = [int(v) for v in sys.argv[1:5]]
dev = torch.device("cuda:0")
qkv = torch.randn(b, l, 3 * heads * d, device=dev).half()
for _ in range(4):
    ops.attention_qkv(qkv, heads, True)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.attention_qkv(qkv, heads, True)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
print(f"attention B{b} L{l} h{heads} d{d}: {ms*1e3:.1f} us  {4.0*b*heads*l*l*d/ms/1e9:.1f} TFLOP/s")
