"""Does the 126 MB L2 keep what a convolution wrote last?  conv (8 x 256 x 256 x 320, tiles run image 0 -> 7) followed by a
GroupNorm over a SUBSET of the images; the GroupNorm alone is timed (CUDA events around it, the conv re-run before every sample).

    python tools/l2_reuse_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops
dev = torch.device("cuda:0")
n, h, w, c = 8, 256, 256, 320
x = torch.randn(n, h, w, c, device=dev).half()
wt = ops.pack_conv_weight(torch.randn(c, c, 3, 3, device=dev) * 0.02, torch.float16)
b = torch.zeros(c, device=dev)
y = torch.empty(n, h, w, c, device=dev, dtype=torch.float16)
st = torch.zeros(n, c, 2, device=dev, dtype=torch.float64)
g, be = torch.ones(c, device=dev), torch.zeros(c, device=dev)


def sample(lo_img, hi_img, reps=8):
    tot = 0.0
    for _ in range(reps + 2):
        st.zero_()
        ops.conv2d([(x, wt, 0, 3, 1)], bias=b, out=y, stats=st)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.group_norm(y[lo_img:hi_img], g, be, cs0=st[lo_img:hi_img])
        e.record(); torch.cuda.synchronize()
        if _ >= 2:
            tot += a.elapsed_time(e)
    return tot / reps * 1e3


for k in (1, 2, 3):
    first, last = sample(0, k), sample(n - k, n)
    mb = 4.0 * k * h * w * c / 1e6
    print(f"GroupNorm of {k} image(s) right after the conv: first-written {first:.1f} us ({mb/first*1e3/1e3:.2f} TB/s)   "
          f"last-written {last:.1f} us ({mb/last*1e3/1e3:.2f} TB/s)")
