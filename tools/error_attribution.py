"""Where does the 16-bit-operand error of one cfg2 / cfg1 network forward come from?  (CPU only, uses the oracle.)

The CPU oracle is run in pure fp32 (the reference) and then with the operands of SELECTED convolutions rounded to fp16
(10-bit mantissa, what the tcgen05 path and the reference's own TF32 GPU path both feed the tensor cores): activations only,
weights only, one resolution at a time, one network section at a time.  The relative L2 deviation of the network output per
selection says which layers would have to run error-compensated (hi + lo operand split) to move the per-step error.

    python tools/error_attribution.py [cfg2|cfg1|cfg2p] > profiles/r2_error_attribution_cfg2.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F

import bench
import dsdiff_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
wl = bench.WORKLOADS[name]
torch.set_num_threads(os.cpu_count() or 1)
sd = O.synth_state_dict(bench.meta_shapes(wl))
cond, x_T = bench.synth_inputs(1, wl["cond_ch"], size, 0)
xc = torch.cat([x_T, cond], 1)
ts = torch.full((1,), 1961, dtype=torch.long)
fwd = O.dsunet_forward if wl["cls"] == "DSUnetModel" else O.unet_forward
h16 = lambda t: t.half().float()
ident = lambda t: t

sel = {"act": lambda p, x: False, "w": lambda p, x: False}


def conv2d(sd_, p, x, stride=1, padding=1, opq=None):
    qa = h16 if sel["act"](p, x) else ident
    qw = h16 if sel["w"](p, x) else ident
    return F.conv2d(qa(x), qw(sd_[p + ".weight"]), sd_.get(p + ".bias"), stride=stride, padding=padding)


O.conv2d = conv2d


def run():
    with torch.no_grad():
        out = fwd(sd, wl["hp"], xc, ts)
    return out[0] if isinstance(out, tuple) else out


def section(p):
    if p.startswith("input_blocks"):
        return "encoders"
    if p.startswith("output_blocks") or p.startswith("out."):
        return "decoder"
    return "middle+disentangle"


t0 = time.time()
ref = run()
res = {"workload": name, "size": size, "fp32_forward_s": time.time() - t0, "selections": {}}
always = lambda p, x: True
never = lambda p, x: False
cases = {"all convs: activations + weights": (always, always),
         "all convs: activations only": (always, never),
         "all convs: weights only": (never, always)}
for r in (size, size // 2, size // 4, size // 8, size // 16, size // 32):
    f = (lambda rr: lambda p, x: x.shape[-1] == rr)(r)
    cases[f"input resolution {r}x{r}: activations + weights"] = (f, f)
for s in ("encoders", "decoder", "middle+disentangle"):
    f = (lambda ss: lambda p, x: section(p) == ss)(s)
    cases[f"section {s}: activations + weights"] = (f, f)
small = lambda p, x: x.shape[-1] <= size // 8
big = lambda p, x: x.shape[-1] > size // 8
cases[f"maps <= {size // 8}x{size // 8} exact, larger maps rounded"] = (big, big)
cases[f"maps <= {size // 4}x{size // 4} exact, larger maps rounded"] = ((lambda p, x: x.shape[-1] > size // 4),) * 2
cases["weights exact everywhere, activations rounded on maps > 32"] = (big, never)
for label, (fa, fw) in cases.items():
    sel["act"], sel["w"] = fa, fw
    t0 = time.time()
    got = run()
    res["selections"][label] = {"rel_l2": O.rel_err(got, ref), "s": round(time.time() - t0, 1)}
    print(label, res["selections"][label], file=sys.stderr, flush=True)
print(json.dumps(res, indent=1))
