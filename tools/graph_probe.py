"""Per-launch device time of small kernels INSIDE a CUDA graph (20 back-to-back nodes, replayed): the fixed cost that the
8x8 .. 32x32 layers of a denoising step pay.  python tools/graph_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops

dev = torch.device("cuda:0")


def probe(name, fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record(); torch.cuda.synchronize()
    print(f"{name:48s} {a.elapsed_time(b) / 5 / reps * 1e3:8.2f} us per launch", flush=True)


def conv(n, h, w, ci, co, k, flags=""):
    x = torch.randn(n, h, w, ci, device=dev).half()
    wt = ops.pack_conv_weight(torch.randn(co, ci, k, k, device=dev) * 0.02, torch.float16)
    b = torch.zeros(co, device=dev)
    out = torch.empty(n, h, w, co, device=dev, dtype=torch.float16)
    kw = {}
    if "s" in flags:
        kw["stats"] = torch.zeros(n, co, 2, device=dev, dtype=torch.float64)
    return lambda: ops.conv2d([(x, wt, 0, k, k // 2)], bias=b, out=out, **kw)


def gn(n, h, w, c):
    x = torch.randn(n, h, w, c, device=dev).half()
    g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    cs = ops.channel_stats(x)
    return lambda: ops.group_norm(x, g, b, cs0=cs)


probe("empty-ish: add_f32 of 1 K floats", (lambda a=torch.zeros(1024, device=dev): (lambda: ops.add_f32(a, a)))())
probe("conv 8x8x8 64->64 k1", conv(8, 8, 8, 64, 64, 1))
probe("conv 8x8x8 960->960 k1", conv(8, 8, 8, 960, 960, 1))
probe("conv 8x8x8 960->960 k3 +stats", conv(8, 8, 8, 960, 960, 3, "s"))
probe("conv 8x16x16 960->960 k3 +stats", conv(8, 16, 16, 960, 960, 3, "s"))
probe("conv 8x32x32 640->640 k3 +stats", conv(8, 32, 32, 640, 640, 3, "s"))
probe("conv 8x32x32 640->1920 k1", conv(8, 32, 32, 640, 1920, 1))
probe("gn 8x8x8x960", gn(8, 8, 8, 960))
probe("gn 8x16x16x960", gn(8, 16, 16, 960))
probe("gn 8x32x32x640", gn(8, 32, 32, 640))
