"""Which STORED activations of a cfg2 (DS-Diff 320-ch) forward have to keep more than 16 bits?  (CPU, oracle pieces.)

Unlike tools/store_attribution.py (which rounds convolution outputs one by one) this walks the network exactly the way the
engine does and rounds at the engine's own storage points: the first convolution's output, conv1 + timestep term, the fused
ResBlock output (skip + conv2, rounded ONCE), Downsample / Upsample convolution outputs, the four-encoder skip mean, the
attention block output; every tensor-core operand is rounded to fp16 (the 10-bit-mantissa floor shared with TF32).
A policy decides which storage points keep the fp32 value (what the engine's hi + lo plane pair holds).

    python tools/store_sim.py > profiles/r2_store_sim_cfg2.txt
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
import bench, dsdiff_oracle as O

wl = bench.WORKLOADS[os.environ.get("WL", "cfg2")]
hp = wl["hp"]
torch.set_num_threads(os.cpu_count())
if os.environ.get("SD") == "module":     # the weights the GPU tests use: module default init (seed 0) + de-zeroed zero_module tensors
    import diffusion_models_dsdiff_b200 as P
    torch.manual_seed(0)
    sd = {k: v.detach().float() for k, v in getattr(P, wl["cls"])(**wl["hp"]).state_dict().items()}
    _g = torch.Generator().manual_seed(1)
    for _k, _v in sd.items():
        if _v.numel() and not _v.any():
            sd[_k] = torch.randn(_v.shape, generator=_g) * 0.02
else:
    sd = O.synth_state_dict(bench.meta_shapes(wl))
cond, x_T = bench.synth_inputs(1, wl["cond_ch"], 256, 0)
xc = torch.cat([x_T, cond], 1)
T_STEP = int(os.environ.get("T_STEP", "1961"))
ts = torch.full((1,), T_STEP, dtype=torch.long)
h16 = lambda t: t.half().float()

POLICY = {"ops": True, "exact": lambda kind, where, res: True, "in_consumers": ("gn", "res", "mean")}


class Pair:
    """The first convolution's output as (exact, rounded): which of its consumers (next GroupNorm / residual add / skip
    mean) read the exact value is part of the policy."""
    def __init__(self, e, r):
        self.e, self.r = e, r

    def pick(self, who):
        return self.e if who in POLICY["in_consumers"] else self.r


def opq():
    return h16 if POLICY["ops"] else None


def store(kind, where, t):
    """kind: in | conv1 | rb | down | up | mean | attn | small ; where: enc | dec | mid ; res = map width"""
    if not POLICY["ops"] or POLICY["exact"](kind, where, t.shape[-1]):
        return t
    return h16(t)


def resblock(p, x, emb, where):
    q = opq()
    x_gn = x
    if isinstance(x, Pair):
        x_gn, x = x.pick("gn"), x.pick("res")
    h = O.silu(O.group_norm(x_gn, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"]))
    h = O.conv2d(sd, p + ".in_layers.2", h, opq=q)
    emb_out = O.linear(sd, p + ".emb_layers.1", O.silu(emb))[:, :, None, None]      # fp32 GEMV in the engine
    if hp.get("use_scale_shift_norm"):      # FiLM (openaimodel.py:276-280): conv1 + bias is stored, the modulation is in the GN pass
        h = store("conv1", where, h)
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = O.silu(O.group_norm(h, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"]) * (1 + scale) + shift)
    else:
        h = store("conv1", where, h + emb_out)
        h = O.silu(O.group_norm(h, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"]))
    h = O.conv2d(sd, p + ".out_layers.3", h, opq=q)
    if p + ".skip_connection.weight" in sd:
        w = sd[p + ".skip_connection.weight"]
        x = O.conv2d(sd, p + ".skip_connection", x, padding=w.shape[-1] // 2, opq=q)
    return store("rb", where, x + h)


def attention(p, x, where):
    q = opq()
    b, c, hh, ww = x.shape
    xs = x.reshape(b, c, -1)
    hn = O.group_norm(xs, sd[p + ".norm.weight"], sd[p + ".norm.bias"])
    qkv = F.conv1d(O._q(q, hn), O._q(q, sd[p + ".qkv.weight"]), sd[p + ".qkv.bias"])
    qkv = store("small", where, qkv)
    heads = O._heads_for(hp, c)
    ch = c // heads
    if hp.get("use_new_attention_order"):
        qq, kk, vv = (t.reshape(b * heads, ch, -1) for t in qkv.chunk(3, dim=1))
    else:
        qq, kk, vv = qkv.reshape(b * heads, ch * 3, -1).split(ch, dim=1)
    scale = ch ** -0.25
    wgt = torch.einsum("bct,bcs->bts", O._q(q, qq * scale), O._q(q, kk * scale)).softmax(-1)
    a = torch.einsum("bts,bcs->bct", O._q(q, wgt), O._q(q, vv)).reshape(b, -1, xs.shape[-1])
    a = store("small", where, a)
    hproj = F.conv1d(O._q(q, a), O._q(q, sd[p + ".proj_out.weight"]), sd[p + ".proj_out.bias"])
    return store("attn", where, (xs + hproj).reshape(b, c, hh, ww))


def sequential(p, x, emb, where):
    q = opq()
    for j in O._sub_indices(sd, p + "."):
        n = f"{p}.{j}"
        if n + ".in_layers.0.weight" in sd:
            x = resblock(n, x, emb, where)
        elif n + ".qkv.weight" in sd:
            x = attention(n, x, where)
        elif n + ".op.weight" in sd:
            x = store("down", where, O.conv2d(sd, n + ".op", x, stride=2, padding=1, opq=q))
        elif n + ".conv.weight" in sd:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = store("up", where, O.conv2d(sd, n + ".conv", x, opq=q))
        elif n + ".weight" in sd:
            y = O.conv2d(sd, n, x, opq=None)                            # fp32 CUDA-core first conv
            x = store("in", where, y)
            if POLICY["ops"] and x is y and len(POLICY["in_consumers"]) < 3:
                x = Pair(y, h16(y))
        else:
            raise KeyError(n)
    return x


def forward():
    q = opq()
    emb = O.time_embed(sd, hp, ts)
    x = xc
    inp = {"": x[:, 0:1], "_a": x[:, 1:2], "_al": x[:, 2:3], "_l": x[:, 3:4]}
    skips, last = {}, {}
    for suffix, h in inp.items():
        hs = []
        for i in O._sub_indices(sd, f"input_blocks{suffix}."):
            h = sequential(f"input_blocks{suffix}.{i}", h, emb, "enc")
            hs.append(h.pick("mean") if isinstance(h, Pair) else h)
        skips[suffix], last[suffix] = hs, h
    h_n = sequential("middle_block", last[""], emb, "mid")
    st = lambda t: store("small", "mid", t)

    def fd(name, v):
        h = O.silu(O.group_norm(v, sd[name + ".conv_1.0.weight"], sd[name + ".conv_1.0.bias"]))
        out = st(O.conv2d(sd, name + ".conv_1.2", h, opq=q) + v)
        h = O.silu(O.group_norm(out, sd[name + ".conv_2.0.weight"], sd[name + ".conv_2.0.bias"]))
        return st(O.conv2d(sd, name + ".conv_2.2", h, padding=0, opq=q))

    h_a, h_al, h_l = last["_a"], last["_al"], last["_l"]
    style = [fd("conv_style", v) for v in (h_a, h_al, h_l)]
    content = [fd("conv_content", v) for v in (h_a, h_al, h_l)]
    anatomy = [fd("conv_anatomy", v) for v in (h_a, h_al)]
    lesion = [fd("conv_lesion", v) for v in (h_al, h_l)]

    def proj(name, lst):
        m = st(torch.mean(torch.stack(lst), dim=0))
        return st(O.conv2d(sd, name + ".1", st(O.se_attention(sd, name + ".0", m)), opq=q))

    h = torch.cat([h_n, proj("share_content_proj", content), proj("style_proj", style), proj("anatomy_proj", anatomy),
                   proj("lesion_proj", lesion)], dim=1)
    h = st(O.conv2d(sd, "all_proj.1", st(O.silu(h)), padding=0, opq=q))
    for i in O._sub_indices(sd, "output_blocks."):
        skip = (skips[""].pop() + skips["_a"].pop() + skips["_al"].pop() + skips["_l"].pop()) / 4
        skip = store("mean", "dec", skip)
        h = sequential(f"output_blocks.{i}", torch.cat([h, skip], dim=1), emb, "dec")
    return O.out_head(sd, h, q)


def run():
    with torch.no_grad():
        return forward()


if __name__ == "__main__":
    POLICY["ops"] = False
    ref = run()
    with torch.no_grad():
        chk = O.dsunet_forward(sd, hp, xc, ts)[0]
    print("walker vs oracle (pure fp32):", f"{O.rel_err(ref, chk):.2e}", flush=True)
    POLICY["ops"] = True
    big = lambda r, m: r >= m
    cases = {
        "operands only (all stores exact)": lambda k, w, r: True,
        "engine today (every store 16-bit)": lambda k, w, r: False,
        "A: rb+in+mean+up exact at 256": lambda k, w, r: r >= 256 and k in ("rb", "in", "mean", "up"),
        "B: A + conv1 exact at 256": lambda k, w, r: r >= 256,
        "C: dec rb + dec conv1 + mean exact at 256": lambda k, w, r: r >= 256 and w == "dec" and k in ("rb", "conv1", "mean", "up"),
        "D: dec rb + mean exact at 256": lambda k, w, r: r >= 256 and w == "dec" and k in ("rb", "mean", "up"),
        "E: dec rb exact at 256": lambda k, w, r: r >= 256 and w == "dec" and k == "rb",
        "F: dec rb + dec conv1 exact at 256": lambda k, w, r: r >= 256 and w == "dec" and k in ("rb", "conv1"),
        "G: F + in exact at 256": lambda k, w, r: r >= 256 and ((w == "dec" and k in ("rb", "conv1")) or k == "in"),
        "H: A at >=128": lambda k, w, r: r >= 128 and k in ("rb", "in", "mean", "up", "down"),
        "I: everything exact at >=128": lambda k, w, r: r >= 128,
        "J: dec (rb, conv1, mean, up) exact at >=128": lambda k, w, r: r >= 128 and w == "dec",
        "K: enc only (rb, in, conv1) exact at 256": lambda k, w, r: r >= 256 and w == "enc",
    }
    G = cases["G: F + in exact at 256"]
    F_ = cases["F: dec rb + dec conv1 exact at 256"]
    cases2 = {
        "L: G, in exact for the skip mean only": (G, ("mean",)),
        "M: G, in exact for GN + residual only": (G, ("gn", "res")),
        "N: G, in exact for GN + mean": (G, ("gn", "mean")),
        "O: G + mean exact": (lambda k, w, r: G(k, w, r) or (r >= 256 and k == "mean"), ("gn", "res", "mean")),
        "P: G + dec (rb, conv1) at 128": (lambda k, w, r: G(k, w, r) or (r >= 128 and w == "dec" and k in ("rb", "conv1")), ("gn", "res", "mean")),
        "Q: G + up at 256": (lambda k, w, r: G(k, w, r) or (r >= 256 and k == "up"), ("gn", "res", "mean")),
    }
    if os.environ.get("BATCH") == "3":
        ALL = ("in", "rb", "conv1", "mean", "up", "down")
        cases = {
            "engine, every store 16-bit": cases["engine today (every store 16-bit)"],
            "operands only": cases["operands only (all stores exact)"],
            "G (shipped default)": G,
            "G + mean": lambda k, w, r: G(k, w, r) or (r >= 256 and k == "mean"),
            "G + enc conv1": lambda k, w, r: G(k, w, r) or (r >= 256 and k == "conv1"),
            "G + enc rb": lambda k, w, r: G(k, w, r) or (r >= 256 and k == "rb"),
            "G + up": lambda k, w, r: G(k, w, r) or (r >= 256 and k == "up"),
            "G kinds at >= 128": lambda k, w, r: r >= 128 and ((w == "dec" and k in ("rb", "conv1")) or k == "in"),
            "everything at 256": lambda k, w, r: r >= 256,
            "everything at >= 128": lambda k, w, r: r >= 128,
        }
    if os.environ.get("BATCH") == "2":
        cases = {"engine today (every store 16-bit)": cases["engine today (every store 16-bit)"]}
        cases.update(cases2)
    sel = os.environ.get("CASES")
    res = {}
    for name, pol in cases.items():
        if sel and name.split(":")[0] not in sel.split(","):
            continue
        POLICY["in_consumers"] = ("gn", "res", "mean")
        if isinstance(pol, tuple):
            pol, POLICY["in_consumers"] = pol
        POLICY["exact"] = pol
        t0 = time.time()
        r = O.rel_err(run(), ref)
        res[name] = r
        print(f"{name:55s} {r:.4e}  ({time.time() - t0:.0f} s)", flush=True)
    print(json.dumps({"t": T_STEP, "rel_l2": res}))
