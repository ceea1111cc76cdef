"""Find the first op whose image-0 result depends on the batch size."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from _util import golden, synth_sd
import diffusion_models_dsdiff_b200 as P
from diffusion_models_dsdiff_b200 import ops

hp, sd = synth_sd("unet_small")
net = P.UNetModel(**hp, act_dtype=torch.float16); net.load_state_dict(sd); net = net.cuda().eval()
g = golden("unet_small")
x, t = g["x"].cuda(), g["t"].cuda()
log = []
names = ["conv2d", "group_norm", "attention_qkv", "conv3x3_in", "conv3x3_out", "channel_stats", "linear_f32", "timestep_embedding", "upsample_nearest2x"]
orig = {n: getattr(ops, n) for n in names}
def wrap(n):
    def f(*a, **k):
        r = orig[n](*a, **k)
        st = k.get("stats")
        log.append((n, r.clone(), None if st is None else st.clone(), tuple(r.shape)))
        return r
    return f
for n in names: setattr(ops, n, wrap(n))
net(x, t); torch.cuda.synchronize(); l2 = list(log); log.clear()
net(x[:1].contiguous(), t[:1]); torch.cuda.synchronize(); l1 = list(log)
print(len(l1), len(l2))
bad = 0
for i, (a, b) in enumerate(zip(l1, l2)):
    same = torch.equal(a[1][:1], b[1][:1])
    ss = True if a[2] is None else bool(((a[2][:1] - b[2][:1]).abs() <= 1e-12 * b[2][:1].abs().clamp_min(1e-30)).all())
    exact = True if a[2] is None else torch.equal(a[2][:1], b[2][:1])
    if not same or not ss:
        print(i, a[0], a[3], b[3], "out equal", same, "stats close", ss, "stats exact", exact,
              "maxdiff", float((a[1][:1].double() - b[1][:1].double()).abs().max()))
        bad += 1
        if bad > 6: break
