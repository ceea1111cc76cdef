"""DiT-B/2 training row kernels on their cfg5 shapes (32 x 1024 tokens x 768), each timed alone with CUDA events; an
ncu --set full target (python tools/row_kernels_one.py [kernel ...])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops, train_ops as T

dev = torch.device("cuda:0")
N, L, C = 32, 1024, 768
dt = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
x, da, dres = rnd(N, L, C).to(dt), rnd(N, L, C).to(dt), rnd(N, L, C).to(dt)
mod = rnd(N, 6 * C) * 0.1
big = rnd(N * L, 4 * C).to(dt)
qkv = rnd(N * L, 3 * C).to(dt)
dsh, dsc = torch.zeros(N, 6 * C, device=dev), torch.zeros(N, 6 * C, device=dev)
cols = {c: torch.zeros(c, device=dev) for c in (C, 3 * C, 4 * C)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

cases = {
    "layernorm": (lambda: ops.layer_norm(x, None, None, 1e-6, shift=mod[:, :C], scale=mod[:, C:2 * C]), 4.0 * x.numel()),
    "layernorm_plain": (lambda: ops.layer_norm(x, None, None, 1e-6), 4.0 * x.numel()),
    "ln_mod_bwd": (lambda: T.ln_mod_bwd(x, da, mod[:, C:2 * C], dres, dsh[:, :C], dsc[:, C:2 * C]), 8.0 * x.numel()),
    "gate_bwd": (lambda: T.gate_bwd(da, x, mod[:, 2 * C:3 * C], dsh[:, 2 * C:3 * C]), 6.0 * x.numel()),
    "gated_residual": (lambda: T.gated_residual(x, da, mod[:, 2 * C:3 * C]), 6.0 * x.numel()),
    "colsum_768": (lambda: T.colsum(x.view(N * L, C), cols[C]), 2.0 * x.numel()),
    "colsum_2304": (lambda: T.colsum(qkv, cols[3 * C]), 2.0 * qkv.numel()),
    "colsum_3072": (lambda: T.colsum(big, cols[4 * C]), 2.0 * big.numel()),
    "gelu": (lambda: T.gelu_fwd(big), 4.0 * big.numel()),
}
sel = sys.argv[1:] or list(cases)
for name in sel:
    fn, nbytes = cases[name]
    fn(); fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        flush.zero_()                      # inputs larger than / evicted from L2 between timed launches
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(e))
    m = sorted(ms)[len(ms) // 2]
    print(f"{name:16s} {m * 1e3:8.1f} us  {nbytes / m / 1e6:8.1f} GB/s")
