"""One attention-backward shape, a few launches (ncu target): python tools/attn_bwd_one.py B L heads d"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import _lib
if os.environ.get("B200DIFF_ABW_DEBUG"):      # timing experiments: a -DB200DIFF_DEBUG_TOGGLES build of the library
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dbg", "libb200diff_dbg.so")
from diffusion_models_dsdiff_b200 import ops, train_ops as T
b, l, heads, d = [int(v) for v in sys.argv[1:5]]
dev = torch.device("cuda:0")
inner = heads * d
qkv = (torch.randn(b, l, 3 * inner, device=dev) * 0.7).bfloat16()
dout = torch.randn(b, l, inner, device=dev).bfloat16()
lse = torch.empty(b, heads, l, device=dev)
att = ops.attention(qkv, qkv, qkv, heads, d, d ** -0.5, q_off=0, k_off=inner, v_off=2 * inner, lse=lse)
for _ in range(3):
    T.attention_bwd(qkv, att, dout, lse, heads, d, d ** -0.5)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    T.attention_bwd(qkv, att, dout, lse, heads, d, d ** -0.5)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 5
print(f"attention_bwd B{b} L{l} h{heads} d{d}: {ms*1e3:.1f} us  {10.0*b*heads*l*l*d/ms/1e9:.1f} algorithmic TFLOP/s "
      f"({14.0*b*heads*l*l*d/ms/1e9:.1f} executed)")
