"""First convolution of a DS-Diff stream (fp32 NCHW -> 16-bit NHWC, CUDA cores, fused GroupNorm statistics), timed alone:

    python tools/conv_in_one.py N H W Cin Cout [lo] [nostats]      prints time, GB/s of the bytes written, and a SHA-1 of the output"""
import hashlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops
n, h, w, ci, co = [int(v) for v in sys.argv[1:6]]
lo = "lo" in sys.argv[6:]
nostats = "nostats" in sys.argv[6:]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
x = torch.randn(n, ci, h, w, device=dev, generator=g)
wt = torch.randn(co, ci, 3, 3, device=dev, generator=g) * 0.2
b = torch.randn(co, device=dev, generator=g) * 0.1
st = torch.zeros(n, co, 2, device=dev, dtype=torch.float64)
run = lambda: ops.conv3x3_in(x, wt, b, torch.float16, stats=None if nostats else st, want_lo=lo)
for _ in range(4):
    run()
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    run()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 10
st.zero_()
out = run()
outs = out if isinstance(out, tuple) else (out,)
sha = hashlib.sha1(b"".join(t.contiguous().view(torch.uint8).cpu().numpy().tobytes() for t in outs) +
                   st.cpu().numpy().tobytes()).hexdigest()[:16]
by = sum(t.numel() * 2 for t in outs)
print(f"conv3x3_in {n}x{ci}x{h}x{w} -> {co}{' +lo' if lo else ''}: {ms:.4f} ms  {by/ms/1e6:.1f} GB/s written  sha {sha}")
