"""Per-(kernel, shape) CUDA-event breakdown of ONE denoising step of a bench workload (diagnosis, not a test).

    python tools/step_breakdown.py [cfg2|cfg1] [batch]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import ops, samplers as S

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device(dev):
    net = getattr(P, wl["cls"])(**wl["hp"]).eval()
if name == "cfg3":      # latent 64x64x3 + 77x768 context
    xc = torch.randn(B, 3, 64, 64, device=dev)
    ctx = torch.randn(B, 77, 768, device=dev)
    fn = lambda x, t, sp: net(x, t, context=ctx, step_ptr=sp)
else:
    xc = torch.randn(B, 1 + wl["cond_ch"], 256, 256, device=dev)
    fn = S._engine_model_fn(net)
t_table = torch.full((50,), 961.0, device=dev)
step_ptr = torch.zeros(1, dtype=torch.int32, device=dev)
fn(xc, t_table, step_ptr); torch.cuda.synchronize()
tm = ops.KernelTimer(); ops.set_timer(tm)
torch.cuda._sleep(int(2e8))
fn(xc, t_table, step_ptr)
ops.set_timer(None)
rows = sorted(tm.by_detail().items(), key=lambda kv: -kv[1]["ms"])
tot = sum(v["ms"] for _, v in rows)
gf = wl["gflop_per_fwd"] * B
print(f"{name} B={B}: {gf/tot:.1f} model TFLOP/s;", end=" ")
print(f"{name} B={B}: {tot:.3f} ms over {sum(v['launches'] for _, v in rows)} timed launches")
for (k, det), v in rows:
    rate = f"{v['flops']/v['ms']/1e9:8.1f} TF/s" if v["flops"] > 0 else f"{v['bytes']/v['ms']/1e6:8.1f} GB/s"
    print(f"{k:16s} {det:44s} x{v['launches']:3d} {v['ms']:8.3f} ms {100*v['ms']/tot:5.1f}%  {rate}")
