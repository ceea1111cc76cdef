"""Per-kernel breakdown of one DiT forward (python tools/dit_breakdown.py [name] [input] [batch])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import ops
name = sys.argv[1] if len(sys.argv) > 1 else "DiT-B/8"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 192
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device(dev):
    net = P.DiT_models[name](input_size=size, in_channels=4).eval()
g = torch.Generator(device=dev).manual_seed(1)
with torch.no_grad():
    for p_ in net.parameters():
        if p_.numel() and not bool(p_.any()):
            p_.copy_(torch.randn(p_.shape, generator=g, device=dev) * 0.02)
net.invalidate()
x = torch.randn(B, 4, size, size, device=dev); t = torch.full((B,), 500.0, device=dev)
net(x, t); torch.cuda.synchronize()
tm = ops.KernelTimer(); ops.set_timer(tm)
torch.cuda._sleep(int(1e8)); net(x, t); ops.set_timer(None)
rows = sorted(tm.by_detail().items(), key=lambda kv: -kv[1]["ms"])
tot = sum(v["ms"] for _, v in rows)
print(f"{name} input {size} B={B}: {tot:.3f} ms in {sum(v['launches'] for _, v in rows)} timed launches")
for (k, det), v in rows[:14]:
    rate = f"{v['flops']/v['ms']/1e9:8.1f} TF/s" if v["flops"] > 0 else f"{v['bytes']/v['ms']/1e6:8.1f} GB/s"
    print(f"{k:14s} {det:40s} x{v['launches']:3d} {v['ms']:8.3f} ms {100*v['ms']/tot:5.1f}%  {rate}")
