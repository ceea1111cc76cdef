"""Diagnosis: grouped vs separate DS-Diff launches vs the fp32 oracle (python tools/group_debug.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import dsdiff_oracle as O
from _util import golden, synth_sd
import diffusion_models_dsdiff_b200 as P

hp, sd = synth_sd("ds_small")
g = golden("ds_small")
for n in (1, 2, 3, 4):
    x = torch.cat([g["x"][:1] * (1.0 + 0.1 * i) for i in range(n)])
    t = g["t"][:1].repeat(n)
    with torch.no_grad():
        ref = O.dsunet_forward(sd, hp, x, t)[0]
    outs = {}
    for label, mp, streams in (("grouped", 1024, True), ("separate", 0, True), ("grouped-1stream", 1024, False)):
        net = P.DSUnetModel(**hp, act_dtype=torch.float16)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        net.group_max_pixels = mp
        net.encoder_streams = streams
        y, _ = net(x.cuda(), t.cuda())
        outs[label] = y.cpu()
    print(f"n={n}: " + "  ".join(f"{k} vs fp32 {O.rel_err(v, ref):.3e}" for k, v in outs.items())
          + f"  grouped vs separate {O.rel_err(outs['grouped'], outs['separate']):.3e}", flush=True)
    for i in range(n):
        print(f"     slice {i}: grouped vs separate {O.rel_err(outs['grouped'][i], outs['separate'][i]):.3e}"
              f"  grouped vs fp32 {O.rel_err(outs['grouped'][i], ref[i]):.3e}  separate vs fp32 {O.rel_err(outs['separate'][i], ref[i]):.3e}")
