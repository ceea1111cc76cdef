"""Throughput of the other BASELINE.json configurations on one B200 (CUDA events, after warm-up; synthetic data,
random-init weights with the zero-initialised layers refilled).  Not the contract bench (bench.py) — evidence that the
same engine serves them.

    python tools/bench_configs.py [cfg1] [cfg2p] [cfg3] [cfg4] [dit]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import samplers as S, parallel as PAR

dev = torch.device("cuda:0")
quiet = open(os.devnull, "w")


def build(name):
    wl = bench.WORKLOADS[name]
    torch.manual_seed(0)
    with torch.device(dev):
        net = getattr(P, wl["cls"])(**wl["hp"]).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for p_ in net.parameters():
            if p_.numel() and not bool(p_.any()):
                p_.copy_(torch.randn(p_.shape, generator=g, device=dev) * 0.02)
    net.invalidate()
    return wl, net


def timed(fn, reps=2, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def timed_graph(fn, reps=5):
    """Device-limited time of one call: captured as a CUDA graph (like the step engine does) and replayed."""
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    return timed(g.replay, reps=reps)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def run_cfg1():
    wl, net = build("cfg1")
    model = S.DDPM(net, timesteps=2000, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
                   conditioning_key="concat", parameterization="v").to(dev)
    for B in (1, 8):
        cond, x_T = [t.to(dev) for t in bench.synth_inputs(B, 3, 256, 0)]
        old = sys.stdout; sys.stdout = quiet
        ms = timed(lambda: S.DDIMSampler(model).sample(50, B, (1, 256, 256), cond, eta=0.0, x_T=x_T, verbose=False))
        sys.stdout = old
        emit(config="cfg1 UNetModel 160-ch (v2-1-cddpm.yaml) DDIM-50 256x256", batch=B, ms_per_batch=ms,
             slices_per_s=B / ms * 1e3, model_tflops=1679.7 * B * 50 / ms)


def run_cfg2p(volume=False):
    wl, net = build("cfg2p")
    wrap = S.DiffusionWrapper(net, "concat")
    for steps in (20, 50):
        d = S.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                        timestep_respacing=str(steps), parameterization="v")
        if not volume:
            B = 64
            cond, x_T = [t.to(dev) for t in bench.synth_inputs(B, 3, 256, 0)]
            ms = timed(lambda: d.ddim_sample_loop(wrap, (B, 1, 256, 256), noise=x_T, model_kwargs={"c_concat": [cond]},
                                                  eta=0.0, progress=False))
            emit(config=f"cfg2' DSUnetModel 96-ch learned-sigma (shipped inference yaml), guided-diffusion DDIM-{steps} 256x256",
                 batch=B, ms_per_batch=ms, slices_per_s=B / ms * 1e3)
        else:
            cond = (torch.randn(155, 3, 240, 240, generator=torch.Generator().manual_seed(3)) * 0.5).clamp_(-1, 1)
            fn = lambda c, x: d.ddim_sample_loop(wrap, tuple(x.shape), noise=x, model_kwargs={"c_concat": [c]}, eta=0.0)
            ms = timed(lambda: PAR.sample_volume(fn, cond, batch_size=64, device=dev), reps=1, warm=1)
            emit(config=f"cfg4 BraTS volume: 155 slices 240x240 (padded to 256), 96-ch DS-Diff, DDIM-{steps}, one GPU",
                 ms_per_volume=ms, slices_per_s=155 / ms * 1e3)


def run_cfg3():
    wl, net = build("cfg3")
    B = 16
    x = torch.randn(B, 3, 64, 64, device=dev)
    ctx = torch.randn(B, 77, 768, device=dev)
    t = torch.full((B,), 500, device=dev)
    ms = timed_graph(lambda: net(x, t, context=ctx))
    emit(config="cfg3 SD-v1 cross-attention UNetModel (v1-inference.yaml, in/out 3 ch) 64x64 latent, one network evaluation",
         batch=B, ms_per_eval=ms, evals_per_s=B / ms * 1e3, model_tflops=803.3 * B / ms)
    ae = P.AutoencoderKL(ddconfig=dict(double_z=True, z_channels=3, resolution=256, in_channels=1, out_ch=1, ch=128,
                                       ch_mult=[1, 2, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0),
                         embed_dim=3).to(dev).eval()
    z = torch.randn(B, 3, 64, 64, device=dev)
    ms = timed_graph(lambda: ae.decode(z))
    emit(config="cfg3 AutoencoderKL (autoencoder_kl_64x64x3.yaml) decode 64x64x3 -> 256x256", batch=B, ms_per_batch=ms,
         slices_per_s=B / ms * 1e3)


def run_dit():
    torch.manual_seed(0)
    with torch.device(dev):
        net = P.DiT_models["DiT-B/8"](input_size=192, in_channels=4).eval()
    g = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for p_ in net.parameters():
            if p_.numel() and not bool(p_.any()):
                p_.copy_(torch.randn(p_.shape, generator=g, device=dev) * 0.02)
    net.invalidate()
    B = 64
    x = torch.randn(B, 4, 192, 192, device=dev)
    t = torch.full((B,), 500.0, device=dev)
    ms = timed_graph(lambda: net(x, t))
    emit(config="DiT-B/8 input 192 in_channels 4 (trainer_use_gaussian_diff_adv.py), one forward", batch=B, ms_per_eval=ms,
         evals_per_s=B / ms * 1e3)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg2p", "cfg3", "cfg4", "dit"]
    with torch.no_grad():
        if "cfg1" in which: run_cfg1()
        if "cfg2p" in which: run_cfg2p()
        if "cfg4" in which: run_cfg2p(volume=True)
        if "cfg3" in which: run_cfg3()
        if "dit" in which: run_dit()
