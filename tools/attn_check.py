"""A/B check of the tcgen05 flash attention against fp32 torch on the GPU + timing."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops

dev = torch.device("cuda:0")
def ref_attn(q, k, v, heads, d, scale):
    b, lq, _ = q.shape; lk = k.shape[1]
    qh = q.float().view(b, lq, heads, d).permute(0, 2, 1, 3)
    kh = k.float().view(b, lk, heads, d).permute(0, 2, 1, 3)
    vh = v.float().view(b, lk, heads, d).permute(0, 2, 1, 3)
    w = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (w @ vh).permute(0, 2, 1, 3).reshape(b, lq, heads * d)

def timeit(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / it

for (b, lq, lk, heads, d) in [(1, 128, 128, 1, 32), (2, 256, 256, 4, 32), (2, 100, 100, 3, 48), (1, 1024, 1024, 2, 32),
                              (2, 64, 64, 30, 32), (1, 200, 77, 8, 64), (1, 256, 300, 2, 160), (1, 1024, 77, 8, 80),
                              (8, 1024, 1024, 20, 32), (8, 256, 256, 30, 32), (8, 64, 64, 30, 32)]:
    for dt in (torch.float16, torch.bfloat16):
        torch.manual_seed(0)
        q = torch.randn(b, lq, heads * d, device=dev).to(dt)
        k = torch.randn(b, lk, heads * d, device=dev).to(dt)
        v = torch.randn(b, lk, heads * d, device=dev).to(dt)
        scale = d ** -0.5
        ref = ref_attn(q, k, v, heads, d, scale)
        out = ops.attention(q, k, v, heads, d, scale)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item()
        line = f"B{b} Lq{lq} Lk{lk} h{heads} d{d} {str(dt)[6:]}: flash err {err:.3e}"
        if d <= 64:
            out2 = ops.attention(q, k, v, heads, d, scale, force_simt=True)
            line += f"  simt err {(out2.float() - ref).abs().max().item():.3e}"
        if dt == torch.float16:
            ms = timeit(lambda: ops.attention(q, k, v, heads, d, scale))
            line += f"  flash {ms*1e3:.1f} us {4.0*b*heads*lq*lk*d/ms/1e9:.1f} TF"
            if d <= 64:
                ms2 = timeit(lambda: ops.attention(q, k, v, heads, d, scale, force_simt=True))
                line += f"  simt {ms2*1e3:.1f} us"
        print(line, flush=True)
