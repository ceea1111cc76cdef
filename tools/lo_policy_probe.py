"""Per-step network-output error of a bench workload at full size with and without the hi + lo planes (GPU vs CPU oracle).

    python tools/lo_policy_probe.py cfg2p [t]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import bench, dsdiff_oracle as O
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200.unet import LoPolicy

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2p"
t = int(sys.argv[2]) if len(sys.argv) > 2 else 961
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
cond, x_T = bench.synth_inputs(1, wl["cond_ch"], 256, 0)
xc = torch.cat([x_T, cond], 1)
ts = torch.tensor([t])
fwd = O.dsunet_forward if wl["cls"] == "DSUnetModel" else O.unet_forward
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    ref = fwd(sd, wl["hp"], xc, ts)
ref = ref[0] if isinstance(ref, tuple) else ref
for label, pol in (("16-bit storage everywhere", None), ("in,dec_rb,dec_conv1", LoPolicy(("in", "dec_rb", "dec_conv1"))),
                   ("dec_rb,dec_conv1", LoPolicy(("dec_rb", "dec_conv1"))),
                   ("everything >= 128x128", LoPolicy(("in", "enc_rb", "dec_rb", "enc_conv1", "dec_conv1", "mean", "up"), 128 * 128))):
    net.lo_policy = pol
    got = net(xc.cuda(), ts.cuda())
    got = (got[0] if isinstance(got, tuple) else got).cpu()
    print(f"{name} t={t} {label:28s} rel L2 vs fp32 oracle {O.rel_err(got, ref):.4e}", flush=True)

# the reference's OWN modules on this GPU (byte-compiled oracle/_ref): TF32 "high" as shipped vs strict fp32
import _ref_import
if _ref_import.available():
    import contextlib, io
    _ref_import.install()
    if wl["cls"] == "DSUnetModel":
        from UNet_DS_Diff.model import DSUnetModel as Ref
    else:
        from ldm.modules.diffusionmodules.openaimodel import UNetModel as Ref
    with contextlib.redirect_stdout(io.StringIO()):
        rn = Ref(**wl["hp"]).eval()
    rn.load_state_dict(sd, strict=True)
    rn = rn.cuda()
    first = lambda o: (o[0] if isinstance(o, tuple) else o).float().cpu()
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.set_float32_matmul_precision("highest")
        strict = first(rn(xc.cuda(), ts.cuda()))
        torch.backends.cudnn.allow_tf32 = True
        torch.set_float32_matmul_precision("high")
        shipped = first(rn(xc.cuda(), ts.cuda()))
    print(f"{name} t={t} reference TF32 as shipped vs its own fp32: {O.rel_err(shipped, strict):.4e}; "
          f"reference fp32 (GPU) vs CPU oracle: {O.rel_err(strict, ref):.2e}")
