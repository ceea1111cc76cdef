"""How much of the per-step error of a cfg2 network forward comes from STORING convolution outputs as 16-bit?  (CPU, oracle.)

The oracle runs in pure fp32 (the reference), then with every convolution's operands rounded to fp16 (what tcgen05 /
TF32 feed the tensor cores) and, additionally, the OUTPUTS of selected convolutions rounded to fp16 (the engine keeps
activations as NHWC fp16 in HBM; the reference's GPU path keeps them fp32).

    python tools/store_attribution.py > profiles/r2_store_attribution_cfg2.txt
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT,"oracle")): sys.path.insert(0,p)
import torch, torch.nn.functional as F
import bench, dsdiff_oracle as O
wl = bench.WORKLOADS["cfg2"]; size=256
torch.set_num_threads(os.cpu_count())
sd = O.synth_state_dict(bench.meta_shapes(wl))
cond, x_T = bench.synth_inputs(1, wl["cond_ch"], size, 0)
xc = torch.cat([x_T, cond], 1); ts = torch.full((1,), 1961, dtype=torch.long)
h16 = lambda t: t.half().float()
mode = {"ops": False, "out": lambda p,x: False}
def conv2d(sd_, p, x, stride=1, padding=1, opq=None):
    q = h16 if mode["ops"] else (lambda t:t)
    y = F.conv2d(q(x), q(sd_[p+".weight"]), sd_.get(p+".bias"), stride=stride, padding=padding)
    return h16(y) if mode["out"](p,x) else y
O.conv2d = conv2d
def run():
    with torch.no_grad():
        out = O.dsunet_forward(sd, wl["hp"], xc, ts)
    return out[0] if isinstance(out, tuple) else out
ref = run()
cases = {
 "operands only": (True, lambda p,x: False),
 "operands + all conv outputs": (True, lambda p,x: not p.startswith("out.")),
 "operands + outputs except conv1 (in_layers.2) at >=128": (True, lambda p,x: not p.startswith("out.") and not (p.endswith("in_layers.2") and x.shape[-1]>=128)),
 "operands + outputs except conv1 at 256": (True, lambda p,x: not p.startswith("out.") and not (p.endswith("in_layers.2") and x.shape[-1]>=256)),
 "operands + outputs except all convs at 256": (True, lambda p,x: not p.startswith("out.") and x.shape[-1]<256),
 "operands + outputs except decoder convs at 256": (True, lambda p,x: not p.startswith("out.") and not (x.shape[-1]>=256 and p.startswith("output_blocks"))),
 "operands + outputs except decoder conv1 at 256": (True, lambda p,x: not p.startswith("out.") and not (x.shape[-1]>=256 and p.startswith("output_blocks") and p.endswith("in_layers.2"))),

 "conv2+skip outputs exact at 256": (True, lambda p,x: not p.startswith("out.") and not (x.shape[-1]>=256 and (p.endswith("out_layers.3") or "skip_connection" in p))),
 "conv2+skip outputs exact at >=128": (True, lambda p,x: not p.startswith("out.") and not (x.shape[-1]>=128 and (p.endswith("out_layers.3") or "skip_connection" in p))),
 "all outputs exact at >=128": (True, lambda p,x: not p.startswith("out.") and x.shape[-1]<128),
 "all outputs exact at >=64": (True, lambda p,x: not p.startswith("out.") and x.shape[-1]<64),
 "conv2+skip+first conv exact at 256": (True, lambda p,x: not p.startswith("out.") and not (x.shape[-1]>=256 and (p.endswith("out_layers.3") or "skip_connection" in p or p.endswith(".0.0")))),
}
for k,(o,f) in cases.items():
    mode["ops"], mode["out"] = o, f
    t0=time.time(); r = O.rel_err(run(), ref)
    print(k, f"{r:.4e}", round(time.time()-t0,1), flush=True)
