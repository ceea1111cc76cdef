"""Bitwise check of a GroupNorm build against another build of the same ABI: prints one SHA-1 per case.

    B200DIFF_LIB=<other .so> python tools/gn_bitcheck.py   vs   python tools/gn_bitcheck.py   -> identical lines"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_models_dsdiff_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
r = lambda *s, scale=1.0: torch.randn(*s, device=dev, generator=g) * scale


def sha(t):
    return hashlib.sha1(t.contiguous().view(torch.uint8).cpu().numpy().tobytes()).hexdigest()[:16]


for dt in (torch.float16, torch.bfloat16):
    n, h, w, c = 3, 40, 24, 320
    x = (r(n, h, w, c) * 3 - 1).to(dt)
    x[0, 0, 0, :8] = -60.0        # exercises the exponent clamp of the SiLU
    x1 = r(n, h, w, 160).to(dt)
    lo = r(n, h, w, c, scale=1e-3).to(dt)
    ga, be = r(c) + 1, r(c)
    ga2, be2 = r(c + 160) + 1, r(c + 160)
    film = r(n, 2 * c, scale=0.3)
    cs, cs1 = ops.channel_stats(x), ops.channel_stats(x1)
    print(dt, "plain      ", sha(ops.group_norm(x, ga, be, cs0=cs)))
    print(dt, "no silu    ", sha(ops.group_norm(x, ga, be, cs0=cs, silu=False)))
    print(dt, "film       ", sha(ops.group_norm(x, ga, be, cs0=cs, film=film)))
    print(dt, "concat     ", sha(ops.group_norm(x, ga2, be2, x1=x1, cs0=cs, cs1=cs1)))
    print(dt, "lo plane   ", sha(ops.group_norm(x, ga, be, cs0=cs, x0_lo=lo)))
    print(dt, "stats pass ", sha(ops.group_norm(x, ga, be)))
