"""Per-kernel timings on one B200 (CUDA events, L2 flushed between iterations)."""
import json, math, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops, _lib as L
import diffusion_models_dsdiff_b200 as P

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timeit(fn, iters=10, warm=3, flush_l2=True):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush_l2: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2]

def conv_case(n, h, w, ci, co, k=3, stride=1, dt=torch.float16):
    x = torch.randn(n, h, w, ci, device=dev).to(dt)
    wt = ops.pack_conv_weight(torch.randn(co, ci, k, k, device=dev) * 0.02, dt)
    b = torch.zeros(co, device=dev)
    out = torch.empty(n, (h + 2*(k//2) - k)//stride + 1, (w + 2*(k//2) - k)//stride + 1, co, device=dev, dtype=dt)
    ms = timeit(lambda: ops.conv2d([(x, wt, 0, k, k//2)], stride=stride, bias=b, out=out))
    fl = 2.0 * out.shape[0]*out.shape[1]*out.shape[2]*co*ci*k*k
    print(json.dumps({"op": "conv", "shape": [n,h,w,ci,co,k,stride], "ms": round(ms,4), "tflops": round(fl/ms/1e9,1)}), flush=True)

def gn_case(n, h, w, c, dt=torch.float16):
    x = torch.randn(n, h, w, c, device=dev).to(dt)
    g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    ms = timeit(lambda: ops.group_norm(x, g, b))
    by = x.numel()*2*2
    print(json.dumps({"op": "gn_silu(stats+apply)", "shape": [n,h,w,c], "ms": round(ms,4), "alg_GBs": round(by/ms/1e6,1)}), flush=True)

def sampler_case(n):
    x = torch.randn(n,1,256,256, device=dev); m = torch.randn_like(x); z = torch.randn_like(x)
    coef = torch.rand(4,16, device=dev)
    ms = timeit(lambda: ops.sampler_update(L.SAMPLER_LDM_DDIM, L.PARAM_V, x, m, z, coef, step=1, x_prev=x, want_x0=False))
    by = x.numel()*4*4
    print(json.dumps({"op": "sampler_update", "B": n, "ms": round(ms,4), "alg_GBs": round(by/ms/1e6,1)}), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "conv"):
        for c in [(1,256,256,320,320), (1,256,256,160,160), (1,128,128,320,320), (1,128,128,640,640), (1,64,64,640,640),
                  (1,64,64,1280,640), (1,32,32,640,640), (8,256,256,320,320), (8,64,64,640,640), (8,32,32,640,640),
                  (8,16,16,960,960), (8,8,8,960,960), (1,256,256,320,160,1), (1,128,128,320,320,3,2)]:
            conv_case(*c)
    if which in ("all", "gn"):
        for c in [(1,256,256,160), (1,256,256,320), (1,128,128,640), (8,256,256,320), (8,32,32,640)]:
            gn_case(*c)
    if which in ("all", "sampler"):
        for n in (1, 64, 512): sampler_case(n)
    if which in ("all", "unet"):
        cfg1 = dict(image_size=32, in_channels=4, model_channels=160, out_channels=1, num_res_blocks=2,
                    attention_resolutions=[16], channel_mult=[1,2,4,4], num_head_channels=32,
                    use_new_attention_order=True, legacy=False)
        net = P.UNetModel(**cfg1).cuda().eval()
        for B in (1, 4):
            x = torch.randn(B,4,256,256, device=dev); t = torch.full((B,), 801, device=dev)
            n0 = P.launch_count(); net(x, t); torch.cuda.synchronize(); nl = P.launch_count() - n0
            ms = timeit(lambda: net(x, t), iters=5)
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                net(x, t)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g, stream=s):
                y = net(x, t)
            msg = timeit(lambda: g.replay(), iters=5)
            print(json.dumps({"op": "UNetModel cfg1 fwd", "B": B, "launches": nl, "eager_ms": round(ms,3), "graph_ms": round(msg,3),
                              "tflops_graph": round(1679.7*B/msg,1)}), flush=True)
