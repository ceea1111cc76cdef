"""One-off: the FULL BASELINE workload for one slice — DS-Diff 320-ch (cfg2) or UNetModel 160-ch (cfg1), 256x256, DDIM-50,
v-parameterisation, T=2000, eta=0 — through the CUDA path (graph engine, hoisted conditioning, forked streams) and through
the CPU oracle, same weights / cond / x_T.  Reports the north_star gates: per-step network-output relative error (the CUDA
network evaluated on the ORACLE's x_t, so only one step's error is measured), final-slice max-abs error, PSNR / MS-SSIM / SUV
deltas.  ~3 min of CPU time per workload on the box's host cores.

    python tools/full_loop_parity.py [cfg2|cfg1] > gpurun_out/full_loop_parity.json
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import bench
import dsdiff_oracle as O
import metrics as M
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import samplers as S

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device(dev):
    net = getattr(P, wl["cls"])(**wl["hp"]).eval()
sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
g = torch.Generator().manual_seed(1)
for k, v in sd.items():
    if v.numel() and not v.any():
        sd[k] = torch.randn(v.shape, generator=g) * 0.02
net.load_state_dict(sd, strict=True)
model = S.DDPM(net, timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
               conditioning_key="concat", parameterization="v", clip_denoised=True).to(dev)
cond, x_T = bench.synth_inputs(1, wl["cond_ch"], 256, 0)
steps = 50
old = sys.stdout; sys.stdout = sys.stderr
out, _ = S.DDIMSampler(model).sample(steps, 1, (1, 256, 256), cond.to(dev), eta=0.0, x_T=x_T.to(dev), verbose=False)
sys.stdout = old
out = out.cpu()

torch.set_num_threads(os.cpu_count() or 1)
sch = O.LdmSchedule(timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2)
tab = O.LdmDdimTables(sch, steps, 0.0)
fwd = O.dsunet_forward if wl["cls"] == "DSUnetModel" else O.unet_forward
x = x_T.clone()
noise = torch.zeros_like(x)
eps_rel = {}
t0 = time.time()
with torch.no_grad():
    for i in range(steps):
        index = steps - 1 - i
        ts = torch.full((1,), int(tab.timesteps[index]), dtype=torch.long)
        xc = torch.cat([x, cond], 1)
        ref = fwd(sd, wl["hp"], xc, ts)
        ref = ref[0] if isinstance(ref, tuple) else ref
        if i in (0, 10, 25, 40, 49):
            got = net(xc.to(dev), ts.to(dev))
            got = (got[0] if isinstance(got, tuple) else got).cpu()
            eps_rel[int(tab.timesteps[index])] = O.rel_err(got, ref)
        x, _ = O.ldm_ddim_step(sch, tab, x, ref, index, noise, "v", True)
cpu_s = time.time() - t0
a, b = out[0, 0].numpy(), x[0, 0].numpy()
yy, xx = np.meshgrid(np.linspace(0, 4 * np.pi, 256), np.linspace(0, 4 * np.pi, 256), indexing="ij")
gt = np.clip(b + 0.15 * np.sin(yy) * np.cos(xx), -1, 1)[None]
suv = lambda v: M.suv_mean(M.pet_to_suv(M.inverse_normalize(v, 0.0, 32000.0), 70.0, 3.7e8, 6586.2, 3600.0))
print(json.dumps({
    "workload": bench.NAMES[name], "slices": 1, "ddim_steps": steps,
    "final_slice_max_abs_err": float(np.abs(a - b).max()), "final_slice_mean_abs_err": float(np.abs(a - b).mean()),
    "per_step_network_output_rel_l2_at_t": eps_rel,
    "psnr_db_cuda_vs_oracle": [M.psnr(gt, a[None]), M.psnr(gt, b[None])],
    "ms_ssim_cuda_vs_oracle": [M.ms_ssim(gt.copy(), a[None].copy()), M.ms_ssim(gt.copy(), b[None].copy())],
    "suv_mean_cuda_vs_oracle": [suv(a), suv(b)],
    "oracle_cpu_seconds": cpu_s, "cores": os.cpu_count()}))
