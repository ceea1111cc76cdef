"""Diagnosis: fast vs generic conv epilogue (run once per env setting; compares with the previous dump)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusion_models_dsdiff_b200 import ops
tag = sys.argv[1]
torch.manual_seed(0)
res = {}
for (n, h, w, ci, co) in [(2, 16, 16, 64, 64), (2, 8, 8, 96, 96), (1, 32, 32, 64, 128), (4, 8, 8, 64, 64)]:
    x = torch.randn(n, h, w, ci).half().cuda()
    wt = (torch.randn(co, ci, 3, 3) / (3 * ci ** 0.5)).cuda()
    b = torch.randn(co).cuda() * 0.1
    rb = torch.randn(n, co).cuda() * 0.1
    r = torch.randn(n, h, w, co).half().cuda()
    st = torch.zeros(n, co, 2, dtype=torch.float64, device="cuda")
    y = ops.conv2d([(x, ops.pack_conv_weight(wt), 0, 3, 1)], bias=b, rowbias=rb, residual=r, stats=st)
    torch.cuda.synchronize()
    yf = y.float()
    exact = torch.stack([yf.double().sum((1, 2)), (yf.double() ** 2).sum((1, 2))], -1)
    print(tag, (n, h, w, ci, co), "stats vs exact-of-stored: max rel", float(((st - exact).abs() / (exact.abs() + 1e-3)).max()))
    res[(n, h, w, ci, co)] = (y.cpu(), st.cpu())
path = "/tmp/epi_dump.pt"
if os.path.exists(path):
    old = torch.load(path)
    for k in res:
        dy = float((res[k][0].float() - old[k][0].float()).abs().max())
        ds = float(((res[k][1] - old[k][1]).abs() / (old[k][1].abs() + 1e-3)).max())
        print("  vs previous dump", k, "max |dy|", dy, "max rel dstats", ds)
torch.save(res, path)
