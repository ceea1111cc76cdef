"""Where does cfg2' (96-ch FiLM DS-Diff) lose accuracy beyond operand + storage rounding?  Diagnosis only: swaps single engine
ops for fp32 torch re-implementations on the GPU and reports the network-output error vs the CPU oracle each time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
import bench, dsdiff_oracle as O
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import ops

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2p"
wl = bench.WORKLOADS[name]
torch.manual_seed(0)
with torch.device("cuda"):
    net = getattr(P, wl["cls"])(**wl["hp"]).eval()
sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
g = torch.Generator().manual_seed(1)
for k, v in sd.items():
    if v.numel() and not v.any():
        sd[k] = torch.randn(v.shape, generator=g) * 0.02
net.load_state_dict(sd, strict=True)
net.lo_policy = None
cond, x_T = bench.synth_inputs(1, wl["cond_ch"], 256, 0)
xc = torch.cat([x_T, cond], 1)
ts = torch.tensor([961])
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    ref = O.dsunet_forward(sd, wl["hp"], xc, ts)[0]
    ref16 = O.dsunet_forward(sd, wl["hp"], xc, ts, opq=lambda t: t.half().float())[0]
print(f"oracle with fp16 operands only: {O.rel_err(ref16, ref):.4e}")


def run(label):
    got = net(xc.cuda(), ts.cuda())[0].cpu()
    print(f"{label:50s} {O.rel_err(got, ref):.4e}", flush=True)


run("engine as shipped (16-bit storage)")
orig_attn, orig_gn = ops.attention_qkv, ops.group_norm


def torch_attn(qkv, heads, new_order):
    b, l, c3 = qkv.shape
    ch = c3 // (3 * heads)
    x = qkv.float()
    if new_order:
        q, k, v = x.chunk(3, dim=-1)
        sp = lambda t: t.reshape(b, l, heads, ch).permute(0, 2, 1, 3)
    else:
        x = x.reshape(b, l, heads, 3 * ch)
        q, k, v = x.split(ch, dim=-1)
        sp = lambda t: t.permute(0, 2, 1, 3)
    q, k, v = sp(q), sp(k), sp(v)
    w = torch.softmax(q @ k.transpose(-1, -2) * (ch ** -0.5), dim=-1)
    return (w @ v).permute(0, 2, 1, 3).reshape(b, l, heads * ch).to(qkv.dtype)


def torch_gn(x0, gamma, beta, x1=None, groups=32, eps=1e-5, silu=True, film=None, **kw):
    if kw.get("preshift") is not None or kw.get("wsets", 1) > 1:
        return orig_gn(x0, gamma, beta, x1, groups=groups, eps=eps, silu=silu, film=film, **kw)
    x = x0 if x1 is None else torch.cat([x0, x1], -1)
    xf = x.float().permute(0, 3, 1, 2)
    y = F.group_norm(xf, groups, gamma, beta, eps)
    if film is not None:
        c = y.shape[1]
        y = y * (1 + film[:, :c, None, None]) + film[:, c:, None, None]
    if silu:
        y = F.silu(y)
    return y.permute(0, 2, 3, 1).contiguous().to(x0.dtype)


ops.attention_qkv = torch_attn
run("attention -> fp32 torch")
ops.attention_qkv = orig_attn
ops.group_norm = torch_gn
net.group_max_pixels = 0
run("GroupNorm(+FiLM)+SiLU -> fp32 torch (ungrouped)")
ops.attention_qkv = torch_attn
run("both")
ops.attention_qkv, ops.group_norm = orig_attn, orig_gn
run("engine, ungrouped launches")
