"""One conv shape, a few launches (ncu --set full target).

    python tools/conv_one.py N H W Cin Cout [k] [flags: s=fused GN statistics, r=residual, b=timestep row bias,
                                                        l=write the low plane (out_lo), L=residual low plane,
                                                        2=stride 2 (Downsample)]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops
n, h, w, ci, co = [int(v) for v in sys.argv[1:6]]
k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
flags = sys.argv[7] if len(sys.argv) > 7 else ""
dev = torch.device("cuda:0")
x = torch.randn(n, h, w, ci, device=dev).half()
wt = ops.pack_conv_weight(torch.randn(co, ci, k, k, device=dev) * 0.02, torch.float16)
b = torch.zeros(co, device=dev)
st = 2 if "2" in flags else 1
out = torch.empty(n, h // st, w // st, co, device=dev, dtype=torch.float16)
kw = {}
if "s" in flags:
    kw["stats"] = torch.zeros(n, co, 2, device=dev, dtype=torch.float64)
if "r" in flags:
    kw["residual"] = torch.randn(n, h, w, co, device=dev).half()
if "b" in flags:
    kw["rowbias"] = torch.randn(n, co, device=dev)
if "l" in flags:
    kw["out_lo"] = torch.empty_like(out)
if "L" in flags:
    kw["residual_lo"] = (torch.randn(n, h, w, co, device=dev) * 1e-3).half()
if k == 2:      # one parity class of a folded Upsample: 2x2 taps, rows / columns {i - 1, i}, dense output of the input's size
    kw["out_hw"] = (h, w)
run = lambda: ops.conv2d([(x, wt, 0, k, (1, 1) if k == 2 else k // 2)], stride=st, bias=b, out=out, **kw)
for _ in range(6):
    run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    run()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
print(f"conv {n}x{h}x{w} {ci}->{co} k{k} [{flags}]: {ms:.4f} ms  {2.0*n*(h//st)*(w//st)*co*ci*k*k/ms/1e9:.1f} TFLOP/s")
