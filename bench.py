#!/usr/bin/env python
"""bench.py — DDIM-50 256x256 slices/sec of the DS-Diff sampling hot path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic input = a full 50-step DDIM sampling
(ldm DDIMSampler, v-parameterisation, T=2000, eta=0, clip on) of ``--batch`` (default 8) 256x256 slices per
GPU through ``DSUnetModel`` (BASELINE.json configs[1]: 320-ch DS-Diff backbone, 3 MRI conditioning channels
concatenated; 64 slices over 8 GPUs = 8 slices per GPU, so the per-GPU work is fixed as N grows: weak
scaling, no data-path collective — slices are independent units).

Rank 0 prints ONE JSON line.  Keys beyond the base contract:
  roofline       dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs / CUDA-event time measured
                 live in this process (one denoising step with per-launch events), vs MEASURED_PEAKS.json
  roofline_aux   the same for the HBM-bound fused kernels (GroupNorm+SiLU, sampler update, ...) and attention
  cpu_baseline   the REFERENCE's own modules (byte-compiled oracle/_ref: its DSDiffModel + DDIMSampler.p_sample_ddim) timed
                 on the host cores on a bounded sample (denoising steps of one slice), extrapolated to slices/s
                 (kind "reference"); the oracle port (oracle/dsdiff_oracle.py, kind "port") only if oracle/_ref is absent
  gpu_eager_reference   the reference's own modules run eagerly on the SAME GPU (TF32 "high" as shipped, and bf16
                 autocast): ms per 8-slice network forward, the slices/s that implies, and each arithmetic's deviation from
                 the reference's strict-fp32 output next to this path's — the bar the hand-written kernels must beat
  other_workloads   short runs of the other BASELINE.json configurations (cfg1 B=1, cfg2' DDIM-20, one cfg4 155-slice
                 volume incl. the gather, cfg3 KL decode)
  e2e            same metric through the public API (DDIMSampler.sample) with HOST (pinned) inputs and a host
                 result: H2D of cond + x_T and D2H of the finished slices inside the timed region
  parity         relative L2 error of the CUDA network output vs the oracle on the cpu_baseline sample

``--impl reference`` times the reference's own CPU implementation of the path — the unmodified reference modules,
byte-compiled into oracle/_ref by oracle/build_ref.py because /root/reference does not exist on the GPU box — on all
host threads; each step is one denoising step (DDIMSampler.p_sample_ddim) of one slice (1/50 of a slice).
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

DDIM_STEPS = 50
WORKLOADS = {
    # configs/v2-1-cddpm-ds-disc.yaml:19-38 (DS-Diff, 320-ch) — BASELINE.json configs[1]
    "cfg2": dict(cls="DSUnetModel", gflop_per_fwd=5544.9, cond_ch=3,
                 hp=dict(image_size=32, in_channels=1, out_channels=1, model_channels=320,
                         attention_resolutions=[32, 16, 8], num_res_blocks=2, channel_mult=[1, 1, 2, 2, 3, 3],
                         num_head_channels=32, use_new_attention_order=True, legacy=False)),
    # configs/v2-1-cddpm.yaml:19-38 (UNetModel, 160-ch) — BASELINE.json configs[0]
    "cfg1": dict(cls="UNetModel", gflop_per_fwd=1679.7, cond_ch=3,
                 hp=dict(image_size=32, in_channels=4, out_channels=1, model_channels=160,
                         attention_resolutions=[16], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                         num_head_channels=32, use_new_attention_order=True, legacy=False)),
    # configs/v2-1-cddpm-ds-disc-openai-diffusion.yaml:44-64 — the 96-ch FiLM / learned-sigma DS-Diff the shipped
    # inference scripts instantiate (SURVEY cfg2'); network-forward workloads only (tools/, tests)
    "cfg2p": dict(cls="DSUnetModel", gflop_per_fwd=506.4, cond_ch=3,
                  hp=dict(image_size=320, in_channels=1, out_channels=2, model_channels=96,
                          attention_resolutions=[32, 16, 8], num_res_blocks=2, channel_mult=[1, 1, 2, 2, 3, 3],
                          num_head_channels=48, use_new_attention_order=True, use_scale_shift_norm=True, legacy=False)),
    # configs/v1-inference.yaml:29-44 — cross-attention UNet of the latent configuration (SURVEY cfg3) at in/out 3 ch
    "cfg3": dict(cls="UNetModel", gflop_per_fwd=803.3, cond_ch=0,
                 hp=dict(image_size=32, in_channels=3, out_channels=3, model_channels=320, attention_resolutions=[4, 2, 1],
                         num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                         transformer_depth=1, context_dim=768, legacy=False)),
}
NAMES = {"cfg2": "DSUnetModel 320-ch (configs/v2-1-cddpm-ds-disc.yaml) 256x256 DDIM-50 v-param T=2000 eta=0",
         "cfg1": "UNetModel 160-ch (configs/v2-1-cddpm.yaml) 256x256 DDIM-50 v-param T=2000 eta=0"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 3; cfg5: 20)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg5"])
    ap.add_argument("--batch", type=int, default=None, help="slices (cfg5: training samples) per GPU per step; default 8 / 32")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--act-dtype", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-pass", action="store_true")
    ap.add_argument("--no-gpu-eager-reference", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tensor_burst": float(p["bf16_tflops"]),
                "tensor": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


def traffic_note(kernel: str):
    """dram bytes/launch from the committed ncu --set full capture (profiles/roofline_traffic.json), else None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of this rank's GPU DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        self.index = vis.split(",")[index] if vis else str(index)
        self.rows, self.proc = [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.index, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.th.join(timeout=3)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synth_inputs(batch, cond_ch, size, rank):
    """SURVEY.md §8d synthetic inputs; per-rank seeds so ranks sample different slices."""
    import torch
    g = torch.Generator().manual_seed(2024 + rank)
    cond = (torch.randn(batch, cond_ch, size, size, generator=g) * 0.5).clamp_(-1, 1)
    x_T = torch.randn(batch, 1, size, size, generator=torch.Generator().manual_seed(2025 + rank))
    return cond, x_T


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the only places bench.py executes oracle/)
# -------------------------------------------------------------------------------------------------
def oracle_step_fn(wl, sd, size):
    """One denoising step of ONE slice with the CPU oracle: network forward + ldm DDIM update."""
    import torch
    import dsdiff_oracle as O
    sch = O.LdmSchedule(timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2)
    tab = O.LdmDdimTables(sch, DDIM_STEPS, 0.0)
    fwd = O.dsunet_forward if wl["cls"] == "DSUnetModel" else O.unet_forward
    cond, x_T = synth_inputs(1, wl["cond_ch"], size, 0)
    index = DDIM_STEPS - 1
    ts = torch.full((1,), int(tab.timesteps[index]), dtype=torch.long)
    noise = torch.zeros_like(x_T)

    def step():
        with torch.no_grad():
            out = fwd(sd, wl["hp"], torch.cat([x_T, cond], 1), ts)
            out = out[0] if isinstance(out, tuple) else out
            xp, _ = O.ldm_ddim_step(sch, tab, x_T, out, index, noise, "v", True)
        return out, xp

    return step, (cond, x_T, ts)


def meta_shapes(wl):
    import torch
    import diffusion_models_dsdiff_b200 as P
    with torch.device("meta"):
        net = getattr(P, wl["cls"])(**wl["hp"])
    return {k: tuple(v.shape) for k, v in net.state_dict().items()}


def reference_available() -> bool:
    import _ref_import
    return _ref_import.available()


def reference_step_fn(wl, sd, size):
    """One denoising step of ONE slice with the REFERENCE ITSELF on the CPU: its trainer class DSDiffModel (ldm DDPM with
    the v-prediction overrides), its own network and its own DDIMSampler.p_sample_ddim (ldm/models/diffusion/ddim.py:188-261)
    at the first (highest-noise) of the 50 DDIM steps.  Returns (step, (cond, x_T, ts))."""
    import torch
    import ref_hosts
    model, sampler = ref_hosts.reference_ddim_host(wl["cls"], wl["hp"], sd, "cpu", size)
    cond, x_T = synth_inputs(1, wl["cond_ch"], size, 0)
    index = DDIM_STEPS - 1
    ts = torch.full((1,), int(sampler.ddim_timesteps[index]), dtype=torch.long)

    def step():                      # what is TIMED: one whole denoising step of the reference
        with torch.no_grad():
            return sampler.p_sample_ddim(x_T, cond, ts, index=index)[0]

    def net_out():                   # the network output of that step (parity reference; not timed)
        with torch.no_grad():
            return model.apply_model(x_T, ts, cond)

    return step, net_out, (cond, x_T, ts)


def cpu_step_fn(wl, sd, size):
    """(step, net_out, inputs, kind): the reference itself when oracle/_ref (or /root/reference) is importable, else the
    oracle port."""
    if reference_available():
        return (*reference_step_fn(wl, sd, size), "reference")
    step, inp = oracle_step_fn(wl, sd, size)
    return step, (lambda: step()[0]), inp, "port"


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import dsdiff_oracle as O
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.synth_state_dict(meta_shapes(wl))
    with contextlib.redirect_stdout(sys.stderr):
        step, _, _, kind = cpu_step_fn(wl, sd, args.size)
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 1.0 / (DDIM_STEPS * dt)
    what = ("the reference's own DDIMSampler.p_sample_ddim over its DSDiffModel + network (unmodified reference, "
            "byte-compiled oracle/_ref)" if kind == "reference" else "oracle port: network forward + DDIM update")
    sample = (f"each step = ONE denoising step ({what}) of ONE {args.size}x{args.size} slice = 1/{DDIM_STEPS} of a slice; "
              f"slices/s = 1/({DDIM_STEPS} x step time)")
    print(json.dumps({
        "impl": "reference", "metric": "DDIM-50 256x256 slices/sec", "value": value, "unit": "slices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": NAMES[args.workload], "slices_per_step": 1.0 / DDIM_STEPS, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": "slices/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
def kernel_pass(model_fn, xc, t_table, **model_kw):
    """One denoising step's network with per-launch CUDA events (the GPU is put to sleep first so the host runs
    ahead and event pairs bracket device execution only)."""
    import torch
    from diffusion_models_dsdiff_b200 import ops
    step_ptr = torch.zeros(1, dtype=torch.int32, device=xc.device)
    model_fn(xc, t_table, step_ptr, **model_kw)
    torch.cuda.synchronize()
    timer = ops.KernelTimer()
    ops.set_timer(timer)
    try:
        torch.cuda._sleep(int(2e8))
        model_fn(xc, t_table, step_ptr, **model_kw)
    finally:
        ops.set_timer(None)
    out = timer.summary()
    # HBM-bound families: the same totals restricted to launches that move >= 32 MB (the rest are launch-latency bound)
    torch.cuda.synchronize()
    for name, fl, by, a, b, det in timer.records:
        if name == "conv_igemm" and det and " k2c" in (" " + det):      # 2x2 sub-pixel launches of a folded Upsample
            out[name]["folded_flops"] = out[name].get("folded_flops", 0.0) + fl
        if fl == 0 and by >= 32e6:
            d = out[name].setdefault("large", {"launches": 0, "ms": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += a.elapsed_time(b)
            d["bytes"] += by
    return out


def gpu_eager_reference(wl, net, cond_d, xT_d, dev, our_value):
    """The reference's OWN modules (oracle/_ref), eager PyTorch on this GPU: one B-slice network forward, three arithmetics —
    strict fp32, TF32 "high" as the reference ships it (inference/inference_2d_main.py:17; cuDNN TF32 is torch's default),
    bf16 autocast.  slices/s = B / (50 x forward time): the sampler's own elementwise launches and host syncs are NOT
    charged to the reference, so this flatters it."""
    import torch
    import dsdiff_oracle as O
    import ref_hosts
    B = xT_d.shape[0]
    ref_net = ref_hosts.reference_network(wl["cls"], wl["hp"])
    ref_net.load_state_dict({k: v.detach().float().cpu() for k, v in net.state_dict().items()}, strict=True)
    ref_net = ref_net.to(dev).eval()
    xc = torch.cat([xT_d, cond_d], 1)
    ts = torch.full((B,), 1961, dtype=torch.long, device=dev)

    def fwd(autocast=False):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            o = ref_net(xc, ts)
        return (o[0] if isinstance(o, tuple) else o).float()

    def modes(name):
        strict = name == "fp32"
        torch.backends.cuda.matmul.allow_tf32 = not strict
        torch.backends.cudnn.allow_tf32 = not strict
        torch.set_float32_matmul_precision("highest" if strict else "high")

    def time_it(autocast):
        fwd(autocast)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            o = fwd(autocast)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 3, o

    out = {"what": "unmodified reference modules (byte-compiled oracle/_ref), eager PyTorch on this GPU, network forward "
                   f"of {B} slices x 3 after one warm-up; slices/s = {B} / ({DDIM_STEPS} x forward)", "batch": B}
    modes("fp32")
    ms32, strict = time_it(False)
    modes("tf32")
    ms_tf, o_tf = time_it(False)
    ms_bf, o_bf = time_it(True)
    modes("fp32")
    with torch.no_grad():
        ours = net(xc, ts)
        ours = (ours[0] if isinstance(ours, tuple) else ours).float()
    for name, ms, o in (("fp32_strict", ms32, None), ("tf32_as_shipped", ms_tf, o_tf), ("bf16_autocast", ms_bf, o_bf)):
        out[name] = {"ms_per_forward": ms, "slices_per_s": B / (DDIM_STEPS * ms / 1e3),
                     "rel_l2_vs_reference_fp32": None if o is None else O.rel_err(o.cpu(), strict.cpu())}
    out["this_path_rel_l2_vs_reference_fp32"] = O.rel_err(ours.cpu(), strict.cpu())
    out["speedup_vs_tf32_as_shipped"] = our_value / out["tf32_as_shipped"]["slices_per_s"]
    out["speedup_vs_bf16_autocast"] = our_value / out["bf16_autocast"]["slices_per_s"]
    del ref_net
    torch.cuda.empty_cache()
    return out


def other_workloads(dev, rank, world):
    """Short device-timed runs of the other BASELINE.json configurations through the same engine (synthetic data, random-init
    weights; one warm-up, then one or two timed runs).  cfg4 shards its 155 slices over the ranks and includes the gather."""
    import torch
    import torch.distributed as dist
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import samplers as S, parallel as PAR

    def build(name):
        w = WORKLOADS[name]
        torch.manual_seed(0)
        with torch.device(dev):
            return w, getattr(P, w["cls"])(**w["hp"]).eval()

    def timed(fn, reps=2, warm=1):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    res = []
    # cfg1 exactly as BASELINE.json states it: UNetModel 160-ch, batch 1, DDIM-50 (per rank; replicas at N > 1)
    w, net = build("cfg1")
    model = S.DDPM(net, timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
                   conditioning_key="concat", parameterization="v").to(dev)
    cond, x_T = [t.to(dev) for t in synth_inputs(1, 3, 256, rank)]
    ms = timed(lambda: S.DDIMSampler(model).sample(DDIM_STEPS, 1, (1, 256, 256), cond, eta=0.0, x_T=x_T, verbose=False))
    res.append({"config": "cfg1 UNetModel 160-ch (configs/v2-1-cddpm.yaml) ldm DDIM-50 256x256, batch 1 per GPU",
                "ms_per_slice": ms, "slices_per_s": world / ms * 1e3, "model_tflops_per_gpu": 1679.7 * 50 / ms})
    del model, net
    # cfg2': the 96-ch learned-sigma DS-Diff the shipped inference yaml runs, guided-diffusion DDIM-20, 64 slices per GPU
    w, net = build("cfg2p")
    wrap = S.DiffusionWrapper(net, "concat")
    d20 = S.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                      timestep_respacing="20", parameterization="v")
    cond, x_T = [t.to(dev) for t in synth_inputs(64, 3, 256, rank)]
    ms = timed(lambda: d20.ddim_sample_loop(wrap, (64, 1, 256, 256), noise=x_T, model_kwargs={"c_concat": [cond]}, eta=0.0))
    res.append({"config": "cfg2' DSUnetModel 96-ch learned-sigma (configs/v2-1-cddpm-ds-disc-openai-diffusion.yaml, what "
                          "inference_config.yaml runs), guided-diffusion DDIM-20 256x256, 64 slices per GPU",
                "ms_per_batch": ms, "slices_per_s": 64 * world / ms * 1e3})
    # cfg4: one BraTS volume, 155 slices of 240x240 padded to 256, sharded over the ranks, gathered on rank 0
    vol = (torch.randn(155, 3, 240, 240, generator=torch.Generator().manual_seed(3)) * 0.5).clamp_(-1, 1)
    fn = lambda c, x: d20.ddim_sample_loop(wrap, tuple(x.shape), noise=x, model_kwargs={"c_concat": [c]}, eta=0.0)
    ms = timed(lambda: PAR.sample_volume(fn, vol, batch_size=64, device=dev), reps=1, warm=1)
    res.append({"config": f"cfg4 BraTS volume: 155 slices 240x240 (padded to 256), 96-ch DS-Diff, DDIM-20 "
                          f"(inference_config_BraTs.yaml defaults), sharded over {world} GPU(s), gather to rank 0 included",
                "ms_per_volume": ms, "slices_per_s": 155 / ms * 1e3})
    del wrap, net
    # cfg3: KL-autoencoder decode 64x64x3 -> 256x256 (configs/autoencoder_kl_64x64x3.yaml)
    ae = P.AutoencoderKL(ddconfig=dict(double_z=True, z_channels=3, resolution=256, in_channels=1, out_ch=1, ch=128,
                                       ch_mult=[1, 2, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0),
                         embed_dim=3).to(dev).eval()
    z = torch.randn(16, 3, 64, 64, device=dev)
    ms = timed(lambda: ae.decode(z), reps=3)
    res.append({"config": "cfg3 AutoencoderKL (configs/autoencoder_kl_64x64x3.yaml) decode 64x64x3 -> 256x256, 16 latents "
                          "per GPU, eager launches", "ms_per_batch": ms, "slices_per_s": 16 * world / ms * 1e3})
    del ae
    torch.cuda.empty_cache()
    # cfg5: DiT-B/2 training step (full line: python bench.py --workload cfg5), 32 samples per GPU, eager launches so the
    # per-block gradient all-reduce runs at N > 1 exactly as in the dedicated workload
    from diffusion_models_dsdiff_b200.dit_train import DiTTrainer
    torch.manual_seed(0)
    with torch.device(dev):
        dit = P.DiT(**CFG5, act_dtype=torch.bfloat16)
    g5 = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p_ in dit.parameters():
            if p_.requires_grad and not p_.detach().any():
                p_.copy_((torch.randn(p_.shape, generator=g5) * 0.02).to(dev))
    gd = S.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                     timestep_respacing="", parameterization="v")
    tr = DiTTrainer(dit, gd, lr=1e-4, act_dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(100 + rank)
    x0 = torch.clamp(torch.randn(32, 1, 64, 64, generator=gen) * 0.5, -1, 1).to(dev)
    cd = torch.clamp(torch.randn(32, 3, 64, 64, generator=gen) * 0.5, -1, 1).to(dev)
    tt = torch.randint(0, 1000, (32,), generator=gen)
    ms = timed(lambda: tr.training_step(x0, tt, cond=[cd], use_graph=False), reps=5, warm=3)
    res.append({"config": "cfg5 " + CFG5_NAME + f"; 32 samples per GPU, bf16, data parallel x{world}",
                "ms_per_step": ms, "samples_per_s": 32 * world / ms * 1e3})
    del tr, dit
    torch.cuda.empty_cache()
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import samplers as S, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    act = torch.float16 if args.act_dtype == "f16" else torch.bfloat16
    B, size = args.batch, args.size

    torch.manual_seed(0)
    with torch.device(dev):
        net = getattr(P, wl["cls"])(**wl["hp"], act_dtype=act)
    net.eval()
    model = S.DDPM(net, timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
                   conditioning_key="concat", parameterization="v", clip_denoised=True).to(dev)
    cond_h, xT_h = synth_inputs(B, wl["cond_ch"], size, rank)
    cond_h, xT_h = cond_h.pin_memory(), xT_h.pin_memory()
    out_h = torch.empty_like(xT_h).pin_memory()
    cond_d, xT_d = cond_h.to(dev), xT_h.to(dev)
    quiet = contextlib.redirect_stdout(sys.stderr)

    def sample_resident():
        smp = S.DDIMSampler(model)
        out, _ = smp.sample(DDIM_STEPS, B, (1, size, size), cond_d, eta=0.0, x_T=xT_d, verbose=False,
                            log_every_t=10 ** 9)
        return out, smp

    from diffusion_models_dsdiff_b200 import parallel as PAR
    gather_ev = []
    out_all_h = torch.empty((B * world, 1, size, size)).pin_memory() if (world > 1 and rank == 0) else None

    def sample_e2e():
        """The call a user makes, host to host: pinned cond + x_T -> device, DDIMSampler.sample, and the finished slices
        back in host memory — at N > 1 on RANK 0, through parallel.gather_slices (NCCL gather over NVLink; the only
        collective of the path, after the loop: SURVEY 8e), like the reference's on_predict_end assembles a volume."""
        c = cond_h.to(dev, non_blocking=True)
        x = xT_h.to(dev, non_blocking=True)
        smp = S.DDIMSampler(model)
        out, _ = smp.sample(DDIM_STEPS, B, (1, size, size), c, eta=0.0, x_T=x, verbose=False, log_every_t=10 ** 9)
        if world > 1:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            full = PAR.gather_slices(out.contiguous(), B * world)
            b.record()
            gather_ev.append((a, b))
            if rank == 0:
                out_all_h.copy_(full, non_blocking=True)
        else:
            out_h.copy_(out, non_blocking=True)
        return out, smp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            _, smp = fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), smp

    with quiet:
        for _ in range(max(args.warmup, 1)):
            sample_resident()
        torch.cuda.synchronize()
        n0 = P.launch_count()
        with ClockSampler(local) as clk:
            ms, smp = timed(sample_resident, args.steps)
        eng = smp.last_engine
        per_step = eng.stats.get("graph_launches_per_step", 0)
        gpu_launches = (P.launch_count() - n0) + per_step * DDIM_STEPS * args.steps
        sample_e2e()
        gather_ev.clear()
        ms_e2e, _ = timed(sample_e2e, args.steps)
    gather_ms = (sum(a.elapsed_time(b) for a, b in gather_ev) / len(gather_ev)) if gather_ev else None
    slices = B * world * args.steps
    value = slices / (ms / 1e3)
    e2e_value = slices / (ms_e2e / 1e3)
    # FLOPs EXECUTED per sampling run: the three conditioning encoders' first conv + GN + conv1 of their first ResBlock do
    # not depend on the timestep and are evaluated once per run instead of once per step (dsunet.prepare_conditioning)
    hoisted = 0.0
    if wl["cls"] == "DSUnetModel" and os.environ.get("B200DIFF_HOIST_COND", "1") != "0":
        c0 = wl["hp"]["model_channels"] * wl["hp"]["channel_mult"][0]
        hoisted = 3 * (2.0 * 256 * 256 * c0 * c0 * 9 + 2.0 * 256 * 256 * c0 * 9) / 1e9
    per_fwd = wl["gflop_per_fwd"]
    flops_per_step = (per_fwd * DDIM_STEPS - hoisted * (DDIM_STEPS - 1)) * 1e9 * (size / 256.0) ** 2 * B

    res = {
        "metric": "DDIM-50 256x256 slices/sec", "value": value, "unit": "slices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.act_dtype + " operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": NAMES[args.workload], "slices_per_gpu_per_step": B, "global_batch": B * world,
                   "ddim_steps": DDIM_STEPS, "image": [size, size], "cond_channels": wl["cond_ch"],
                   "parallelism": f"slice-sharded x{world}, no data-path collective",
                   "l2": "activations per layer exceed the 126 MB L2 (inputs larger than L2; no explicit flush)",
                   "cuda_graph": True, "graph_kernels_per_denoise_step": per_step,
                   "activation_storage": ("16-bit NHWC; hi + lo plane pairs (32 bits) for " +
                                          ",".join(sorted(net.lo_policy.kinds)) + f" on maps >= {net.lo_policy.min_pixels} "
                                          "pixels (unet.LoPolicy, B200DIFF_LO_PLANES)") if getattr(net, "lo_policy", None)
                                         else "16-bit NHWC everywhere",
                   "loop_invariant_hoisting": ("conditioning encoders' first conv + first ResBlock GN/conv1 evaluated once "
                                               f"per sampling run ({hoisted:.1f} of {per_fwd:.1f} GFLOP per slice-step); "
                                               "model_tflops_per_gpu counts executed FLOPs") if hoisted else None},
        "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": int(cond_h.numel() * 4 + xT_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel() * 4 * (world if world > 1 else 1)),
                "ms_per_step": ms_e2e / args.steps,
                "gather": None if world == 1 else {
                    "what": "parallel.gather_slices: NCCL gather of every rank's finished slices to rank 0 (inside the "
                            "e2e timed region), then ONE device->pinned-host copy of the whole batch on rank 0",
                    "ms_per_step_this_rank": gather_ms, "bytes_gathered_per_step": int(out_h.numel() * 4 * world)}},
        "gpu_launches": int(gpu_launches),
        "clocks": clk.summary(),
        "model_tflops_per_gpu": flops_per_step * args.steps / (ms / 1e3) / 1e12,
    }

    pk = peaks()
    if rank == 0 and not args.no_kernel_pass:
        with quiet:
            xc = torch.cat([xT_d, cond_d], 1)
            t_table = torch.full((DDIM_STEPS,), 1961.0, device=dev)
            # the step as the sampler runs it: loop-invariant conditioning prefix hoisted out (DSUnetModel.prepare_conditioning)
            mkw = {}
            prep = getattr(net, "prepare_conditioning", None)
            if prep is not None and prep([cond_d]):
                mkw = {"cond_cached": True}
            ks = kernel_pass(model._model_fn(), xc, t_table, **mkw)
            # sampler update at the bench batch (HBM-bound; B*256*256 fp32 elements)
            timer = ops.KernelTimer()
            ops.set_timer(timer)
            coef = torch.rand(DDIM_STEPS, 16, device=dev)
            mo, z = torch.randn_like(xT_d), torch.randn_like(xT_d)
            torch.cuda._sleep(int(2e7))
            for _ in range(5):
                ops.sampler_update(0, 2, xT_d.clone(), mo, z, coef, step=1, want_x0=True)
            ops.set_timer(None)
            ks.update(timer.summary())
            # the same kernel at 512 slices (134 MB of state): at the bench batch it is a 2 MB launch-latency-bound node of
            # the step graph, so its HBM fraction is reported at a size where bandwidth is what limits it
            big = torch.randn(512, 1, size, size, device=dev)
            mo2, z2, xp2, x02 = torch.randn_like(big), torch.randn_like(big), torch.empty_like(big), torch.empty_like(big)
            ops.sampler_update(0, 2, big, mo2, z2, coef, step=1, x_prev=xp2, pred_x0=x02)
            timer = ops.KernelTimer()
            ops.set_timer(timer)
            torch.cuda._sleep(int(2e7))
            for _ in range(5):
                ops.sampler_update(0, 2, big, mo2, z2, coef, step=1, x_prev=xp2, pred_x0=x02)
            ops.set_timer(None)
            ks["sampler_update_512_slices"] = timer.summary()["sampler_update"]
            del big, mo2, z2, xp2, x02
        total_ms = sum(v["ms"] for k, v in ks.items() if not k.startswith("sampler_update"))
        aux = []
        for name, v in sorted(ks.items(), key=lambda kv: -kv[1]["ms"]):
            tensor = v["flops"] > 0 and name in ("conv_igemm", "attention")
            ach = (v["flops"] / v["ms"] / 1e9) if tensor else (v["bytes"] / v["ms"] / 1e6)
            peak = pk["tensor"] if tensor else pk["hbm"]
            entry = {"kernel": name, "bound": "tensor" if tensor else "hbm", "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu --set full capture
                     # (bytes; null when no capture exists); which launch and its algorithmic bytes: traffic_detail
                     "traffic": (traffic_note(name) or {}).get("dram_bytes_per_launch"),
                     "traffic_detail": traffic_note(name),
                     "launches_per_denoise_step": v["launches"], "ms_per_denoise_step": v["ms"],
                     "share_of_step": (v["ms"] / total_ms) if not name.startswith("sampler_update") else None,
                     "peak_source": pk["source"] + (" (sustained)" if tensor else " (copy)")}
            if not tensor and v.get("large"):
                lg = v["large"]
                entry["large_launches"] = {"min_bytes": 32e6, "launches": lg["launches"],
                                           "achieved": lg["bytes"] / lg["ms"] / 1e6, "frac": lg["bytes"] / lg["ms"] / 1e6 / peak,
                                           "share_of_bytes": lg["bytes"] / v["bytes"]}
            aux.append(entry)
        dom = next(e for e in aux if e["kernel"] == "conv_igemm")
        res["roofline"] = {k: dom[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_detail")}
        # executed vs algorithmic: the five Upsample layers run as 2x2 sub-pixel convolutions (16 of their 36 multiply-adds)
        folded = ks["conv_igemm"].get("folded_flops", 0.0)
        res["roofline"].update({"kernel": "conv_igemm_kernel (tcgen05 implicit GEMM)", "peak_source": dom["peak_source"],
                                "share_of_step": dom["share_of_step"],
                                "achieved_executed": (ks["conv_igemm"]["flops"] - folded * 5.0 / 9.0) / ks["conv_igemm"]["ms"] / 1e9,
                                "how": "sum of algorithmic conv FLOPs of one denoising step / sum of per-launch "
                                       "CUDA-event durations on the launching stream"})
        res["roofline_aux"] = aux
        res["kernel_pass_ms_per_denoise_step"] = total_ms

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import dsdiff_oracle as O
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
        with quiet:
            step, net_out, (c1, x1, ts), kind = cpu_step_fn(wl, sd, size)
            ref_out = net_out()                      # warm-up; also the parity reference below
            ref_out = ref_out[0] if isinstance(ref_out, tuple) else ref_out
            n_cpu = 3                                # bounded sample: ~10-30 s of CPU work on the box's host cores
            t0 = time.perf_counter()
            for _ in range(n_cpu):
                step()
            dt = (time.perf_counter() - t0) / n_cpu
        with quiet, torch.no_grad():
            got = net(torch.cat([x1, c1], 1).to(dev), ts.to(dev))
            got = (got[0] if isinstance(got, tuple) else got).cpu()
        what = ("the reference's own DDIMSampler.p_sample_ddim over its DSDiffModel + network, unmodified reference "
                "byte-compiled into oracle/_ref" if kind == "reference" else "oracle port: network forward + DDIM update")
        res["cpu_baseline"] = {
            "value": 1.0 / (DDIM_STEPS * dt), "unit": "slices/s", "cores": cores, "kind": kind,
            "sample": f"{n_cpu} denoising steps ({what}) of one {size}x{size} slice after one warm-up forward, "
                      f"{dt:.2f} s each, x{DDIM_STEPS} extrapolated to a slice"}
        res["parity"] = {"net_out_rel_l2_vs_reference_fp32_cpu": O.rel_err(got, ref_out), "reference_kind": kind,
                         "note": "same weights / input as the cpu_baseline sample; 16-bit operands vs pure fp32; the gate "
                                 "(tests/test_gpu_full_loop.py) is 1.25 x the reference's own TF32 deviation below"}

    if rank == 0 and world == 1 and not args.no_gpu_eager_reference and reference_available():
        with quiet:
            res["gpu_eager_reference"] = gpu_eager_reference(wl, net, cond_d, xT_d, dev, value)
    if not args.no_other_workloads and args.workload == "cfg2":
        del model, net
        torch.cuda.empty_cache()
        with quiet:
            ow = other_workloads(dev, rank, world)
        if rank == 0:
            res["other_workloads"] = ow
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 5: DiT-B/2 conditional diffusion TRAINING step, bf16 operands, gradient all-reduce over NVLink
# ---------------------------------------------------------------------------------------------------------------------
CFG5 = dict(input_size=64, patch_size=2, in_channels=4, hidden_size=768, depth=12, num_heads=12, learn_sigma=True)
CFG5_NAME = ("DiT-B/2 (UNet_DS_Diff/DiT_models.py) conditional diffusion training step: 64x64x(1+3 cond) input -> 1024 tokens, "
             "learned sigma, v-parameterisation, GaussianDiffusion.training_losses (Charbonnier + VB), AdamW, DDP all-reduce")


def cfg5_flops(hp, batch):
    """Algorithmic FLOPs of one training step (forward + backward = 3 x forward matmul FLOPs; attention backward 2.5 x)."""
    d, depth, heads = hp["hidden_size"], hp["depth"], hp["num_heads"]
    tokens = (hp["input_size"] // hp["patch_size"]) ** 2
    lin = 2.0 * batch * tokens * 12 * d * d * depth
    attn = 4.0 * batch * heads * tokens * tokens * (d // heads) * depth
    return {"linear": 3 * lin, "attention_fwd": attn, "attention_bwd": 2.5 * attn}


def run_cfg5(args):
    import torch
    import torch.distributed as dist
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import samplers as S, ops
    from diffusion_models_dsdiff_b200.dit_train import DiTTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or 32
    act = torch.bfloat16                                  # cfg5 is stated in bf16
    torch.manual_seed(0)
    with torch.device(dev):
        net = P.DiT(**CFG5, act_dtype=act)
    # adaLN-Zero / zero-init heads would make most gradients exactly zero: de-zero like every other synthetic workload
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p_ in net.parameters():
            if p_.requires_grad and not p_.detach().any():
                p_.copy_((torch.randn(p_.shape, generator=g) * 0.02).to(dev))
    gd = S.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                     timestep_respacing="", parameterization="v")
    tr = DiTTrainer(net, gd, lr=1e-4, act_dtype=act)
    s = CFG5["input_size"]
    gen = torch.Generator().manual_seed(100 + rank)
    x0_h = torch.clamp(torch.randn(B, 1, s, s, generator=gen) * 0.5, -1, 1).pin_memory()
    cond_h = torch.clamp(torch.randn(B, 3, s, s, generator=gen) * 0.5, -1, 1).pin_memory()
    x0_d, cond_d = x0_h.to(dev), cond_h.to(dev)
    t = torch.randint(0, 1000, (B,), generator=gen)

    def step_resident():
        return tr.training_step(x0_d, t, cond=[cond_d])

    def step_e2e():
        x0, cd = x0_h.to(dev, non_blocking=True), cond_h.to(dev, non_blocking=True)
        out = tr.training_step(x0, t, cond=[cd])
        return float(out["loss"].item())                 # device -> host read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            r = fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), r

    steps = max(args.steps, 1)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    n0, r0 = P.launch_count(), tr.replayed_kernels
    with ClockSampler(local) as clk:
        ms, out = timed(step_resident, steps)
    launches = (P.launch_count() - n0) + (tr.replayed_kernels - r0)      # eager launches + kernels replayed from the step graph
    step_e2e()
    ms_e2e, loss = timed(step_e2e, steps)
    samples = B * world * steps
    fl = cfg5_flops(CFG5, B)
    pk = peaks()
    res = {
        "metric": "DiT-B/2 diffusion training samples/sec", "value": samples / (ms / 1e3), "unit": "samples/s",
        "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16 operands, f32 accumulate, f32 master weights / gradients / AdamW",
        "data": "synthetic", "gpu_launches": launches,
        "config": {"workload": CFG5_NAME, "samples_per_gpu_per_step": B, "global_batch": B * world, "tokens_per_sample": 1024,
                   "parameters": int(tr.flat_p.numel()), "optimizer": "AdamW lr 1e-4 (torch defaults)",
                   "cuda_graph": bool(tr.use_graph and world == 1),
                   "parallelism": f"data parallel x{world}: replicated weights, per-block gradient buckets all-reduced over NCCL "
                                  "as the backward produces them",
                   "l2": "activations per layer (50-200 MB) exceed the 126 MB L2 across a step; no explicit flush"},
        "e2e": {"value": samples / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": int(x0_h.numel() * 4 + cond_h.numel() * 4), "d2h_bytes_per_step": 4,
                "loss": loss},
        "clocks": clk.summary(),
        "model_tflops_per_gpu": sum(fl.values()) / (ms / steps / 1e3) / 1e12,
    }
    if not args.no_kernel_pass:
        # per-kernel CUDA-event pass over ONE step (events on the launching stream)
        tm = ops.KernelTimer()
        ops.set_timer(tm)
        torch.cuda._sleep(int(2e8))
        tr.training_step(x0_d, t, cond=[cond_d], use_graph=False)       # eager: per-launch events need individual launches
        ops.set_timer(None)
        summ = tm.summary()
        tot = sum(v["ms"] for v in summ.values())
        gemm = summ.get("conv_igemm", {"ms": 0.0, "flops": 0.0, "launches": 0})
        res["kernel_pass_ms_per_step"] = tot
        res["roofline"] = {"bound": "tensor", "achieved": gemm["flops"] / max(gemm["ms"], 1e-9) / 1e9, "peak": pk["tensor"],
                           "unit": "TFLOP/s", "frac": gemm["flops"] / max(gemm["ms"], 1e-9) / 1e9 / pk["tensor"],
                           "traffic": None, "kernel": "conv_igemm_kernel (forward, dgrad and split-K wgrad GEMMs)",
                           "share_of_step": gemm["ms"] / max(tot, 1e-9), "peak_source": pk["source"],
                           "how": "sum of algorithmic GEMM FLOPs of one training step / sum of per-launch CUDA-event durations"}
        aux = []
        for name, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
            if name == "conv_igemm":
                continue
            tensor = v["flops"] > 0
            ach = (v["flops"] / v["ms"] / 1e9) if tensor else (v["bytes"] / v["ms"] / 1e6)
            aux.append({"kernel": name, "bound": "tensor" if tensor else "hbm", "achieved": ach,
                        "peak": pk["tensor"] if tensor else pk["hbm"], "unit": "TFLOP/s" if tensor else "GB/s",
                        "frac": ach / (pk["tensor"] if tensor else pk["hbm"]), "launches_per_step": v["launches"],
                        "ms_per_step": v["ms"], "share_of_step": v["ms"] / max(tot, 1e-9)})
        res["roofline_aux"] = aux
    if rank == 0 and world == 1 and not args.no_gpu_eager_reference and reference_available():
        res["gpu_eager_reference"] = cfg5_gpu_eager_reference(B, dev, res["value"])
    # gradient all-reduce alone: the whole fp32 gradient arena, NCCL ring / NVLS
    if world > 1:
        gbytes = tr.flat_g.numel() * 4
        for _ in range(2):
            dist.all_reduce(tr.flat_g)
        ms_ar, _ = timed(lambda: dist.all_reduce(tr.flat_g), 5)
        ms_ar /= 5
        res["allreduce"] = {"bytes": gbytes, "ms_alone": ms_ar, "algbw_GBps": gbytes / ms_ar / 1e6,
                            "busbw_GBps": gbytes / ms_ar / 1e6 * 2 * (world - 1) / world,
                            "note": "inside a step the per-block buckets are issued asynchronously while the backward of the "
                                    "earlier blocks still runs; exposed time = step time at N GPUs - step time at 1 GPU"}
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cfg5_gpu_eager_reference(B, dev, our_value):
    """The UNMODIFIED reference DiT + GaussianDiffusion.training_losses + autograd + torch.optim.AdamW on this GPU (eager
    PyTorch; bf16 autocast as BASELINE cfg5 states; TF32 'high' for the fp32 variant).  timm is absent and unpinned: its
    Attention is restated twice — the classic softmax(q k^T) v and the fused F.scaled_dot_product_attention current timm
    releases dispatch to — and both are reported."""
    import torch
    import torch.nn.functional as F
    import _ref_import
    _ref_import.install()
    import make_golden_ext as MGE
    MGE._install_timm_restatement()
    import sys as _sys
    vt = _sys.modules["timm.models.vision_transformer"]
    classic = vt.Attention

    class FusedAttention(classic):
        def forward(self, x):
            b, n, c = x.shape
            qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
            q, k, v = qkv.unbind(0)
            x = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
            return self.proj(x)

    import training_project.utils.gaussian_diffusion as GD
    from training_project.utils.script_util import create_gaussian_diffusion
    import UNet_DS_Diff.DiT_models as RD
    GD.get_heatmap = lambda m: m
    out = {"what": "unmodified reference modules (byte-compiled oracle/_ref) + torch autograd + torch.optim.AdamW on this GPU, "
                   f"{B} samples per step, 3 timed steps after 2 warm-up steps", "batch": B}
    s = CFG5["input_size"]
    x0 = torch.clamp(torch.randn(B, 1, s, s, device=dev) * 0.5, -1, 1)
    cond = [torch.clamp(torch.randn(B, 1, s, s, device=dev) * 0.5, -1, 1) for _ in range(3)]
    kw = {"F_Data1": cond[0], "F_Data2": cond[1], "S_Data1": cond[2]}
    diff = create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                     timestep_respacing="", parameterization="v")
    for attn_name, attn_cls in (("sdpa_attention", FusedAttention), ("classic_attention", classic)):
        RD.Attention = attn_cls
        for mode in ("bf16_autocast", "tf32"):
            torch.manual_seed(0)
            net = RD.DiT(**CFG5).to(dev)
            with torch.no_grad():
                for p_ in net.parameters():
                    if p_.requires_grad and not p_.detach().any():
                        p_.normal_(0, 0.02)
            opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
            model = lambda x, t, c_concat=None: net(torch.cat([x] + c_concat, 1), t)
            torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
            torch.set_float32_matmul_precision("high" if mode == "tf32" else "highest")

            def step():
                opt.zero_grad(set_to_none=True)
                t = torch.randint(0, 1000, (B,), device=dev)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16_autocast"):
                    terms = diff.training_losses(model, x0, t, model_kwargs=kw)
                    loss = terms["loss"].mean()
                loss.backward()
                opt.step()

            try:
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    step()
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 3
                out[f"{attn_name}.{mode}"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                                              "speedup_of_this_path": our_value / (B / ms * 1e3)}
            except Exception as e:      # e.g. out of memory for the materialised fp32 score matrices
                out[f"{attn_name}.{mode}"] = {"error": f"{type(e).__name__}: {str(e)[:120]}"}
            del net, opt
            torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    return out


def run_reference_cfg5(args):
    """The reference's own DiT + training_losses + backward + AdamW on the host cores (bounded sample: 2 samples per step)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "compiled reference (oracle/_ref) not present"}))
        return
    import _ref_import
    _ref_import.install()
    import make_golden_ext as MGE
    MGE._install_timm_restatement()
    import training_project.utils.gaussian_diffusion as GD
    from training_project.utils.script_util import create_gaussian_diffusion
    from UNet_DS_Diff.DiT_models import DiT
    GD.get_heatmap = lambda m: m
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = DiT(**CFG5)
    diff = create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                     timestep_respacing="", parameterization="v")
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    b, s = 2, CFG5["input_size"]
    x0 = torch.clamp(torch.randn(b, 1, s, s) * 0.5, -1, 1)
    cond = [torch.clamp(torch.randn(b, 1, s, s) * 0.5, -1, 1) for _ in range(3)]
    model = lambda x, t, c_concat=None: net(torch.cat([x] + c_concat, 1), t)

    def step():
        opt.zero_grad()
        terms = diff.training_losses(model, x0, torch.randint(0, 1000, (b,)),
                                     model_kwargs={"F_Data1": cond[0], "F_Data2": cond[1], "S_Data1": cond[2]})
        terms["loss"].mean().backward()
        opt.step()

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    k = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    v = b / dt
    sample = f"{k} full training steps of {b} samples (the reference's DiT-B/2 + training_losses + backward + AdamW, fp32, CPU)"
    print(json.dumps({"impl": "reference", "metric": "DiT-B/2 diffusion training samples/sec", "value": v, "unit": "samples/s",
                      "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": CFG5_NAME}, "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores,
                                                                         "kind": "reference", "sample": sample},
                      "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    args = parse()
    if args.steps is None:
        args.steps = 20 if args.workload == "cfg5" and args.impl != "reference" else 3
    if args.workload == "cfg5":
        (run_reference_cfg5 if args.impl == "reference" else run_cfg5)(args)
        return
    if args.batch is None:
        args.batch = 8
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
