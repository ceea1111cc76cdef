"""GPU parity of the DiT TRAINING step (BASELINE cfg5): every training kernel against a plain PyTorch fp32
restatement of the same op, and the whole step (loss terms, all parameter gradients, parameters after two AdamW updates)
against fixtures produced by the UNMODIFIED reference's DiT + GaussianDiffusion.training_losses + loss.backward() +
torch.optim.AdamW (oracle/make_golden_train_dit.py).

Tolerances: 16-bit tensor-core operands and 16-bit stored activation gradients put a gradient tensor at 1e-3..1e-2 relative L2
(fp16) and ~8x that for bf16 (2^-8 vs 2^-11 rounding), exactly like the forward parity tests; fp32 kernels are held to 1e-5."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import dsdiff_oracle as O
import training_oracle as TO
from _util import GOLD, golden, manifest, maxabs, synth_sd

pytestmark = pytest.mark.gpu
DTYPES = [torch.float16, torch.bfloat16]
EPS16 = {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}


def _T():
    from diffusion_models_dsdiff_b200 import train_ops
    return train_ops


def rel(a, b):
    return O.rel_err(a.float().cpu(), b.float().cpu())


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


# -------------------------------------------------------------------------------------------------
# op level
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DTYPES)
def test_transpose_colsum_gelu(dt):
    T = _T()
    x = rnd(1, 200, 136).to(dt)
    assert torch.equal(T.transpose16(x.cuda()).cpu(), x.t().contiguous())
    out = torch.empty(136, device="cuda")
    T.colsum(x.cuda(), out)
    assert (out.cpu() - x.float().sum(0)).abs().max() < 1e-3
    pre, du = rnd(2, 64, 256, scale=2.0).to(dt), rnd(3, 64, 256).to(dt)
    p32 = pre.float().requires_grad_(True)
    y = F.gelu(p32, approximate="tanh")
    y.backward(du.float())
    assert (T.gelu_fwd(pre.cuda()).float().cpu() - y.detach()).abs().max() < 2 * EPS16[dt] * max(1.0, float(y.abs().max()))
    got = T.gelu_bwd(pre.cuda(), du.cuda()).float().cpu()
    assert (got - p32.grad).abs().max() < 2 * EPS16[dt] * max(1.0, float(p32.grad.abs().max()))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("c", [128, 768, 1152])
def test_gate_and_ln_modulate_backward(dt, c):
    T = _T()
    n, l = 3, 37
    h, f, dh = rnd(4, n, l, c).to(dt), rnd(5, n, l, c).to(dt), rnd(6, n, l, c).to(dt)
    mod = rnd(7, n, 6 * c, scale=0.3)
    gate = mod[:, 2 * c:3 * c]
    ref = h.float() + gate[:, None] * f.float()
    got = T.gated_residual(h.cuda(), f.cuda(), mod.cuda()[:, 2 * c:3 * c]).float().cpu()
    assert (got - ref).abs().max() < 2 * EPS16[dt] * float(ref.abs().max())
    dmod = torch.zeros(n, 6 * c, device="cuda")
    df = T.gate_bwd(dh.cuda(), f.cuda(), mod.cuda()[:, 2 * c:3 * c], dmod[:, 2 * c:3 * c]).float().cpu()
    assert (df - gate[:, None] * dh.float()).abs().max() < 2 * EPS16[dt] * float(dh.abs().max())
    want = (dh.float() * f.float()).sum(1)
    assert rel(dmod[:, 2 * c:3 * c], want) < 1e-5 and float(dmod[:, :2 * c].abs().max()) == 0.0
    # a = LN(x) (1 + scale) + shift
    x32 = (h.float() * 2 + 0.5).to(dt).float().requires_grad_(True)
    sh, sc = mod[:, :c].clone().requires_grad_(True), mod[:, c:2 * c].clone().requires_grad_(True)
    a = F.layer_norm(x32, (c,), eps=1e-6) * (1 + sc[:, None]) + sh[:, None]
    da, dres = rnd(8, n, l, c).to(dt), rnd(9, n, l, c).to(dt)
    a.backward(da.float())
    dmod.zero_()
    dx = T.ln_mod_bwd(x32.detach().to(dt).cuda(), da.cuda(), mod.cuda()[:, c:2 * c], dres.cuda(), dmod[:, :c],
                      dmod[:, c:2 * c]).float().cpu()
    ref_dx = x32.grad + dres.float()
    assert (dx - ref_dx).abs().max() < 3 * EPS16[dt] * float(ref_dx.abs().max())
    assert rel(dmod[:, :c], sh.grad) < 1e-5 and rel(dmod[:, c:2 * c], sc.grad) < 1e-4


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2, 64, 4, 32), (1, 200, 2, 64), (2, 320, 3, 64)])
def test_attention_backward_matches_autograd(dt, shape):
    """Forward log-sum-exp + backward kernel vs torch autograd through the fp32 restatement of timm Attention's core."""
    from diffusion_models_dsdiff_b200 import ops
    T = _T()
    b, l, heads, d = shape
    inner = heads * d
    qkv = rnd(10, b, l, 3 * inner, scale=0.7).to(dt)
    dout = rnd(11, b, l, inner).to(dt)
    scale = d ** -0.5
    q32 = qkv.float().requires_grad_(True)
    qh, kh, vh = (q32[..., i * inner:(i + 1) * inner].view(b, l, heads, d).permute(0, 2, 1, 3) for i in range(3))
    s = qh @ kh.transpose(-1, -2) * scale
    att_ref = (s.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(b, l, inner)
    att_ref.backward(dout.float())
    lse = torch.empty(b, heads, l, device="cuda")
    att = ops.attention(qkv.cuda(), qkv.cuda(), qkv.cuda(), heads, d, scale, q_off=0, k_off=inner, v_off=2 * inner, lse=lse)
    assert (att.float().cpu() - att_ref.detach()).abs().max() < 4 * EPS16[dt] * float(att_ref.abs().max())
    assert (lse.cpu() - torch.logsumexp(s.detach(), -1)).abs().max() < 2e-3
    dqkv = T.attention_bwd(qkv.cuda(), att, dout.cuda(), lse, heads, d, scale)
    for i, nm in enumerate("qkv"):
        got, want = dqkv[..., i * inner:(i + 1) * inner], q32.grad[..., i * inner:(i + 1) * inner]
        assert rel(got, want) < 6 * EPS16[dt], nm


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("t_o_i", [(192, 384, 128), (4096, 768, 768), (2048, 16, 768), (1024, 128, 16),
                                   (520, 192, 320), (8192, 2304, 768), (40, 64, 64)])
def test_linear_wgrad_on_tensor_cores(dt, t_o_i):
    """MN-major wgrad_tn kernel (O, I multiples of 64: ragged 128 x 256 tiles, token counts that are not a multiple of the
    64-token chunk, K splits) and the transposing fallback (16-wide operands), with the bias gradient."""
    T = _T()
    t, o, i = t_o_i
    dy, x = rnd(12, t, o).to(dt), rnd(13, t, i).to(dt)
    db = torch.full((o,), 7.0, device="cuda")
    got = T.linear_wgrad(dy.cuda(), x.cuda(), db)
    want = dy.float().t() @ x.float()
    assert got.dtype == torch.float32 and tuple(got.shape) == (o, i)
    assert rel(got, want) < 1e-5          # operands are exact 16-bit values; fp32 accumulation / split-K order only
    assert rel(db, dy.float().sum(0)) < 1e-5
    out = torch.zeros(o, i, device="cuda")
    assert T.linear_wgrad(dy.cuda(), x.cuda(), out=out).data_ptr() == out.data_ptr() and rel(out, want) < 1e-5


@pytest.mark.parametrize("param,learn_sigma,rescaled", [("v", True, False), ("eps", True, True), ("x0", False, False)])
def test_diffusion_loss_and_its_gradient(param, learn_sigma, rescaled):
    """b200_diffusion_loss vs the training oracle (pinned to the reference's training_losses) and autograd through it."""
    T = _T()
    n, c, s = 4, 1, 16
    sch = O.GdSchedule(steps=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True)
    x0 = torch.clamp(rnd(14, n, c, s, s) * 0.6, -1, 1)
    x0[0, 0, :2] = -1.0
    x0[0, 0, 2:4] = 1.0
    nz = rnd(15, n, c, s, s)
    t = torch.tensor([0, 3, 500, 999])
    out = rnd(16, n, 2 * c if learn_sigma else c, s, s, scale=0.8).requires_grad_(True)
    terms = TO.training_losses(sch, lambda xc, tm: out, x0, t, nz, [], parameterization=param, learn_sigma=learn_sigma,
                               rescaled_mse=rescaled)
    terms["loss"].mean().backward()
    from diffusion_models_dsdiff_b200 import samplers as S
    from diffusion_models_dsdiff_b200.dit_train import DiTTrainer, _PARAM
    gd = S.create_gaussian_diffusion(steps=1000, learn_sigma=learn_sigma, rescale_timesteps=True,
                                     rescale_learned_sigmas=rescaled, parameterization=param,
                                     predict_xstart=param == "x0")
    tr = DiTTrainer.__new__(DiTTrainer)
    tr.diffusion, tr.rescaled_mse = gd, rescaled
    tr.flat_p = torch.zeros(1, device="cuda")
    coef = tr._coef_rows(t)
    x_t = T.q_sample(x0.cuda(), nz.cuda(), coef)
    assert maxabs(x_t.cpu(), TO.q_sample(sch, x0, t, nz)) < 1e-6
    got_terms, d_out = T.diffusion_loss(out.detach().cuda(), x0.cuda(), x_t, nz.cuda(), coef, _PARAM[param], learn_sigma)
    assert rel(got_terms[:, 0], terms["mse"].detach()) < 1e-5
    if learn_sigma:
        assert rel(got_terms[:, 1], terms["vb"].detach()) < 2e-4
    assert rel(d_out, out.grad) < 2e-4


def test_small_fp32_linears_and_adamw():
    T = _T()
    n, o, i = 5, 300, 72
    x, w, dy = rnd(17, n, i), rnd(18, o, i, scale=0.2), rnd(19, n, o)
    x32, w32 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    b32 = torch.zeros(o, requires_grad=True)
    (F.linear(F.silu(x32), w32, b32) * dy).sum().backward()
    dW, db = torch.empty(o, i, device="cuda"), torch.empty(o, device="cuda")
    T.linear_f32_wgrad(dy.cuda(), x.cuda(), dW, db, in_silu=True)
    assert rel(dW, w32.grad) < 1e-5 and rel(db, b32.grad) < 1e-5
    dx = T.linear_f32_dgrad(dy.cuda(), w.cuda(), x_pre=x.cuda())
    assert rel(dx, x32.grad) < 1e-5
    # AdamW: three steps against torch.optim.AdamW
    p = rnd(20, 1000).requires_grad_(True)
    opt = torch.optim.AdamW([p], lr=3e-3)
    pc = p.detach().clone().cuda()
    m, v = torch.zeros_like(pc), torch.zeros_like(pc)
    for step in range(1, 4):
        g = rnd(20 + step, 1000)
        p.grad = g.clone()
        opt.step()
        T.adamw(pc, (g * 2).cuda(), m, v, 3e-3, step=step, grad_scale=0.5)
    assert (pc.cpu() - p.detach()).abs().max() < 1e-6


# -------------------------------------------------------------------------------------------------
# the whole step against the reference's own training step
# -------------------------------------------------------------------------------------------------
def _trainer(name, dt, lr):
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import samplers as S
    from diffusion_models_dsdiff_b200.dit_train import DiTTrainer
    man = manifest()[name]
    g = golden(name)
    hp, sd = synth_sd(name)
    net = P.DiT(**hp, act_dtype=dt)
    sd["pos_embed"] = g["pos_embed"]
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    gd = S.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                     rescale_learned_sigmas=man["rescaled"], timestep_respacing="",
                                     parameterization=man["parameterization"])
    return g, net, DiTTrainer(net, gd, lr=lr, act_dtype=dt)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("name", ["train_dit_v", "train_dit_eps_rescaled"])
def test_dit_training_step_matches_the_reference_training_step(name, dt):
    g, net, tr = _trainer(name, dt, float(g_lr(name)))
    names = open(os.path.join(GOLD, name + ".names.txt")).read().split()
    cond = [c for c in g["cond"]]
    out = tr.training_step(g["x0"], g["t"], cond=cond, noise=g["noise"], update=False)
    tol = 8e-3 if dt == torch.float16 else 5e-2
    assert abs(float(out["loss"]) - float(g["total"])) < tol * abs(float(g["total"]))
    assert rel(tr.last_terms[:, 0], g["mse"]) < tol and rel(tr.last_terms[:, 1], g["vb"]) < tol
    grads = tr.grads()
    gtol = 2e-2 if dt == torch.float16 else 1.2e-1
    for nm, ref in zip(names, g["grad_norms"].tolist()):
        if ref < 0:
            continue
        got = float(grads[nm].double().norm())
        assert abs(got - ref) < gtol * ref, (nm, got, ref)
    for k in g:
        if k.startswith("g:"):
            assert rel(grads[k[2:]].reshape(g[k].shape), g[k]) < gtol, (k, rel(grads[k[2:]].reshape(g[k].shape), g[k]))


def g_lr(name):
    return golden(name)["lr"]


@pytest.mark.parametrize("name", ["train_dit_v", "train_dit_eps_rescaled"])
def test_two_optimizer_steps_reach_the_reference_parameters(name):
    """AdamW's first steps move every weight by ~lr regardless of the gradient's magnitude (sign-like update), so the
    parameters after two steps are a sharp check of the whole chain: a wrong gradient SIGN anywhere shows up as 2 lr."""
    g, net, tr = _trainer(name, torch.float16, float(g_lr(name)))
    cond = [c for c in g["cond"]]
    before = {k: p.detach().clone() for k, p in net.named_parameters()}
    tr.training_step(g["x0"], g["t"], cond=cond, noise=g["noise"])
    out2 = tr.training_step(g["x0"], g["t2"], cond=cond, noise=g["noise2"])
    assert abs(float(out2["loss"]) - float(g["loss2"].mean())) < 2e-2 * abs(float(g["loss2"].mean()))
    lr = float(g["lr"])
    params = dict(net.named_parameters())
    for k in g:
        if k.startswith("p2:"):
            got, want, was = params[k[3:]].detach().cpu().reshape(g[k].shape), g[k], before[k[3:]].cpu().reshape(g[k].shape)
            if k.endswith("attn.qkv.bias"):
                # the K third of the qkv bias has an exactly zero gradient (adding a constant to every key's score leaves
                # the softmax unchanged): Adam then normalises pure roundoff noise into +-lr steps on both sides
                d3 = got.numel() // 3
                keep = torch.cat([torch.arange(0, d3), torch.arange(2 * d3, 3 * d3)])
                got, want, was = got[keep], want[keep], was[keep]
            moved = (want - was).abs().mean()
            assert float(moved) > 0.2 * lr, k                                     # the reference did move it
            # elements whose gradient is near zero may flip their Adam direction under 16-bit noise: mean, not max
            assert float((got - want).abs().mean()) < 0.12 * float(moved), (k, float((got - want).abs().mean()), float(moved))
    # the sampling path sees the updated weights (packed copies were invalidated)
    x = torch.cat([g["x0"]] + cond, 1).cuda()
    y = net(x, torch.tensor([3.0, 500.0, 900.0], device="cuda"))
    assert torch.isfinite(y).all()


def test_graph_replayed_training_steps_equal_eager_steps():
    """The captured-and-replayed step (static input buffers, optimizer scalars read from device memory) walks the same
    parameter trajectory as eager launches: five steps with changing t / noise, the last three replayed."""
    name = "train_dit_v"
    runs = []
    for use_graph in (False, True):
        g, net, tr = _trainer(name, torch.bfloat16, 1e-3)
        tr.use_graph = use_graph
        cond = [c for c in g["cond"]]
        losses = []
        for k in range(5):
            noise = rnd(200 + k, *g["noise"].shape)
            t = torch.tensor([(17 * k + 3) % 1000, (400 + 91 * k) % 1000, 999 - 55 * k])
            losses.append(float(tr.training_step(g["x0"], t, cond=cond, noise=noise)["loss"]))
        assert (len(tr._graphs) == 1) == use_graph
        assert tr.step_count == 5
        runs.append((losses, tr.flat_p.detach().clone()))
    (l0, p0), (l1, p1) = runs
    assert max(abs(a - b) / abs(a) for a, b in zip(l0, l1)) < 2e-3, (l0, l1)
    # fp32 atomics (column sums) add in a run-dependent order, and Adam turns near-zero gradients into +-lr steps:
    # mean, not max
    assert float((p0 - p1).abs().mean()) < 0.05 * 1e-3


def test_dit_b2_training_step_at_the_baseline_size_vs_oracle_autograd():
    """BASELINE cfg5 at its real size — DiT-B/2 (12 x 768, 12 heads of 64), 64 x 64 x (1 + 3) input = 1024 tokens, bf16 —
    on two samples: loss terms and gradients against torch autograd through the CPU oracle's restated forward + objective
    (which tests/test_training_oracle.py pins to the reference's own training step on the small fixtures)."""
    import bench
    import diffusion_models_dsdiff_b200 as P
    from diffusion_models_dsdiff_b200 import samplers as S
    from diffusion_models_dsdiff_b200.dit_train import DiTTrainer
    hp = dict(bench.CFG5)
    torch.manual_seed(0)
    net = P.DiT(**hp, act_dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p_ in net.parameters():
            if p_.requires_grad and not p_.detach().any():
                p_.copy_(torch.randn(p_.shape, generator=g) * 0.02)       # adaLN-Zero / zero heads would zero most gradients
    sd = {k: v.detach().clone().float() for k, v in net.state_dict().items()}
    net = net.cuda()
    gd = S.create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                     timestep_respacing="", parameterization="v")
    tr = DiTTrainer(net, gd, lr=1e-4, act_dtype=torch.bfloat16)
    n, s = 2, hp["input_size"]
    x0 = torch.clamp(rnd(301, n, 1, s, s) * 0.5, -1, 1)
    cond = [torch.clamp(rnd(302 + i, n, 1, s, s) * 0.5, -1, 1) for i in range(3)]
    noise = rnd(310, n, 1, s, s)
    t = torch.tensor([37, 812])
    out = tr.training_step(x0, t, cond=cond, noise=noise, update=False)
    grads = tr.grads()
    # oracle: same weights, fp32, autograd
    sdg = {k: (v.clone().requires_grad_(True) if k != "pos_embed" else v) for k, v in sd.items()}
    sch = O.GdSchedule(steps=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True)
    ohp = dict(patch_size=hp["patch_size"], hidden_size=hp["hidden_size"], depth=hp["depth"], num_heads=hp["num_heads"],
               in_channels=hp["in_channels"], learn_sigma=True)
    terms = TO.training_losses(sch, lambda xc, tm: O.dit_forward(sdg, ohp, xc, tm), x0, t, noise, cond, parameterization="v",
                               learn_sigma=True)
    total = terms["loss"].mean()
    total.backward()
    assert abs(float(out["loss"]) - float(total)) < 3e-2 * abs(float(total))
    assert rel(tr.last_terms[:, 0], terms["mse"].detach()) < 3e-2
    worst = 0.0
    for k in ("final_layer.linear.weight", "blocks.11.mlp.fc2.weight", "blocks.11.attn.qkv.weight", "blocks.6.attn.proj.weight",
              "blocks.6.mlp.fc1.bias", "blocks.0.attn.qkv.weight", "blocks.0.adaLN_modulation.1.weight", "x_embedder.proj.weight",
              "t_embedder.mlp.0.weight", "final_layer.adaLN_modulation.1.bias"):
        r = rel(grads[k].reshape(sdg[k].grad.shape), sdg[k].grad)
        worst = max(worst, r)
        assert r < 1.5e-1, (k, r)        # bf16 activations + activation gradients through 12 blocks
    print(f"\n[cfg5 full size] loss {float(out['loss']):.5f} vs oracle {float(total):.5f}; worst gradient rel L2 {worst:.3e}")
