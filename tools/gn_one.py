"""One GroupNorm+SiLU shape, a few launches (ncu --set full target): python tools/gn_one.py N H W C [lo]

    lo: the source comes with a low plane (hi + lo pair, unet.LoPolicy): 6 instead of 4 algorithmic bytes per element"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops
n, h, w, c = [int(v) for v in sys.argv[1:5]]
dev = torch.device("cuda:0")
x = torch.randn(n, h, w, c, device=dev).half()
g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
cs = ops.channel_stats(x)
lo = (torch.randn(n, h, w, c, device=dev) * 1e-3).half() if len(sys.argv) > 5 and sys.argv[5] == "lo" else None
for _ in range(4):
    ops.group_norm(x, g, b, cs0=cs, x0_lo=lo)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.group_norm(x, g, b, cs0=cs, x0_lo=lo)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
per = 6.0 if lo is not None else 4.0
print(f"gn+silu {n}x{h}x{w}x{c}{' +lo' if lo is not None else ''}: {ms:.4f} ms  {per*x.numel()/ms/1e6:.1f} GB/s "
      f"(algorithmic: {per:.0f} B per element)")
