"""cfg1 (UNetModel 160-ch) DDIM-50 at batch 1, timed: python tools/cfg1_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contextlib, io, torch, bench
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import samplers as S
dev = torch.device("cuda:0")
wl = bench.WORKLOADS["cfg1"]
torch.manual_seed(0)
with torch.device(dev):
    net = P.UNetModel(**wl["hp"]).eval()
model = S.DDPM(net, timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2, conditioning_key="concat",
               parameterization="v").to(dev)
cond, x_T = [t.to(dev) for t in bench.synth_inputs(1, 3, 256, 0)]
def run():
    with contextlib.redirect_stdout(io.StringIO()):
        return S.DDIMSampler(model).sample(50, 1, (1, 256, 256), cond, eta=0.0, x_T=x_T, verbose=False)
run(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    out, smp = run(), None
b.record(); torch.cuda.synchronize()
print(f"cfg1 B=1 DDIM-50: {a.elapsed_time(b)/3:.1f} ms per slice; fold={os.environ.get('B200DIFF_NO_UPSAMPLE_FOLD','0')!='1'}")
