"""Print the parity numbers behind the GPU model tests (not a test: diagnosis)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import dsdiff_oracle as O
from _util import golden, maxabs, synth_sd
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import samplers as S

def net_for(case, cls, dtype=torch.float16):
    hp, sd = synth_sd(case)
    net = cls(**hp, act_dtype=dtype); net.load_state_dict(sd, strict=True)
    return hp, sd, net.cuda().eval()

for case in ("ds_small", "ds_small_2ch"):
    hp, sd, net = net_for("ds_small", P.DSUnetModel)
    g = golden(case)
    y, _ = net(g["x"].cuda(), g["t"].cuda())
    yq, _ = O.dsunet_forward(sd, hp, g["x"], g["t"], opq=lambda v: v.half().float())
    print(f"{case}: vs fp32 golden {O.rel_err(y.cpu(), g['y']):.3e}; oracle(fp16 operands) vs golden {O.rel_err(yq, g['y']):.3e}; "
          f"gpu vs oracle(fp16 operands) {O.rel_err(y.cpu(), yq):.3e}", flush=True)

for mode, eta, tag in (("ddpm", 0.0, "ls"), ("ddim", 0.0, "ls"), ("ddpm", 0.0, "fixed")):
    g = golden(f"v34_{mode}_{tag}_eta{eta}")
    hp, sd, net = net_for("ds_film2" if tag == "ls" else "ds_small", P.DSUnetModel)
    wrap = S.DiffusionWrapper(net, "concat")
    d = S.create_gaussian_diffusion(steps=1000, learn_sigma=(tag == "ls"), timestep_respacing="5",
                                    rescale_timesteps=True, parameterization="v")
    rec = []
    fn = d.ddim_sample_loop if mode == "ddim" else d.p_sample_loop
    kw = {"eta": eta} if mode == "ddim" else {}
    out = fn(wrap, tuple(g["x_T"].shape), noise=g["x_T"].cuda(), clip_denoised=True,
             model_kwargs={"c_concat": [g["cond"].cuda()]}, noises=list(g["noises"]), record=rec, **kw)
    for i in range(g["S"]):
        print(f"{mode}-{tag} step {i}: "
              f"x0 maxabs {maxabs(rec[i]['pred_x0'].cpu(), g['pred_x0'][i]):.3e} x_prev maxabs {maxabs(rec[i]['x_prev'].cpu(), g['x_inter'][i]):.3e} "
              f"|x_prev|max {g['x_inter'][i].abs().max():.2f}", flush=True)
