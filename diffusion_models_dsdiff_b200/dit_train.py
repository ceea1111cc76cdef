"""DiT diffusion TRAINING step on libb200diff (BASELINE config 5).

What the reference does per step (trainers/trainer_use_gaussian_diff_adv.py:345-382 with the ``ViT_config`` denoiser
``DiT_B_*`` behind ``DiffusionWrapper`` 'concat', :74-77; ``configure_optimizers`` :306-320 = ``torch.optim.AdamW(params,
lr)``; DDP gradient all-reduce precedent ``Disc_diff/guided_diffusion/train_util.py:113-120``):

    t            = schedule_sampler.sample(B)                                     (uniform)
    x_t          = q_sample(x_start, t, noise)                                    gaussian_diffusion.py:199-215
    out          = DiT(cat[x_t, cond], scaled t)                                   UNet_DS_Diff/DiT_models.py:234-254
    loss         = mean_n( Charbonnier(target(v | eps | x0), out_mean) + vb(out_var) )   gaussian_diffusion.py:824-983
    loss.backward() ; all-reduce(grads) ; AdamW.step()

Here: the forward runs the same tcgen05 GEMM / flash-attention / LayerNorm kernels as sampling but keeps the activations
the backward needs; the backward is hand-derived (no autograd): per DiTBlock two gate / LayerNorm-modulate backward passes,
GELU backward, the attention backward kernel, and for every Linear a dgrad GEMM (transposed packed weight) and a wgrad GEMM
(fp32, split-K over all tokens) on the same tensor-core kernel.  Parameters, gradients and the AdamW moments live in flat
fp32 arenas (the module's ``nn.Parameter``s are views of the parameter arena, so ``state_dict`` / sampling see every
update); the gradient arena is all-reduced in per-block buckets as soon as a block's backward has finished, overlapping
NCCL with the remaining backward, and one fused AdamW kernel updates the whole arena.

Scope: ``y`` (class labels) is not part of this trainer (the reference's DiT configuration conditions through the
concatenated images only); head width must be a multiple of 16 (DiT-S/B/L: 64).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from . import train_ops as T
from .dit import DiT, _f
from .transformer import linear_tokens

Tensor = torch.Tensor
_PARAM = {"eps": 0, "x0": 1, "v": 2}


class _Lin:
    """One Linear's packed weights for this step: ``w`` [1, O, I] (forward) and ``wt`` [1, I, O] (dgrad)."""
    __slots__ = ("w", "wt", "b", "gw", "gb", "o", "i")


def arena_layout(named_shapes, depth: int):
    """Flat-arena placement of the trainable parameters (in ``named_parameters`` order, 32-byte aligned slots) and the
    gradient all-reduce buckets: one per DiTBlock — its parameters are contiguous — plus everything before / after the blocks.
    ``named_shapes``: [(name, numel)].  Returns (slots {name: slice}, total, block_ranges, head_range, tail_range)."""
    slots: Dict[str, slice] = {}
    off = 0
    for k, numel in named_shapes:
        slots[k] = slice(off, off + numel)
        off += -(-numel // 8) * 8
    total = off
    block_ranges: List[slice] = []
    for i in range(depth):
        ks = [k for k in slots if k.startswith(f"blocks.{i}.")]
        block_ranges.append(slice(min(slots[k].start for k in ks), max(slots[k].stop for k in ks)))
    return slots, total, block_ranges, slice(0, block_ranges[0].start), slice(block_ranges[-1].stop, total)


def reduce_bucket(flat_g: Tensor, rng: slice, group, handles: list) -> None:
    """Asynchronous SUM all-reduce of one contiguous gradient bucket (issued as soon as the bucket is complete; the caller
    waits on the handles before the optimizer step; AdamW divides by the world size)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1 and rng.stop > rng.start:
        handles.append(dist.all_reduce(flat_g[rng], group=group, async_op=True))


class DiTTrainer:
    def __init__(self, model: DiT, diffusion, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, parameterization: Optional[str] = None, rescaled_mse: bool = False,
                 process_group=None, act_dtype: torch.dtype = torch.bfloat16):
        if not isinstance(model, DiT):
            raise L.B200Error("DiTTrainer drives the package's DiT")
        dev = model.pos_embed.device
        if dev.type != "cuda":
            raise L.B200Error("the B200 engine has no CPU path: move the model to a CUDA device")
        d, h = model.hidden_size, model.num_heads
        if (d // h) % 16 or (d // h) not in (32, 64):
            raise NotImplementedError("DiT training: head width must be 32 or 64")
        self.model, self.diffusion = model, diffusion
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        # the target follows the reference's precedence: parameterization "v" overrides model_mean_type
        # (gaussian_diffusion.py:338-342, 960-970)
        if parameterization is None:
            parameterization = "v" if getattr(diffusion, "parameterization", "eps") == "v" else diffusion.model_mean_type
        self.param = _PARAM[parameterization]
        self.rescaled_mse = rescaled_mse or getattr(diffusion, "loss_type", None) == "rescaled_mse"
        self.group = process_group
        self.dtype = act_dtype
        self.step_count = 0
        # ---- flat arenas; nn.Parameters become views of the parameter arena ----
        # y_embedder is never reached when y is None (the reference's configuration): torch leaves its .grad None and
        # AdamW skips it — so it stays out of the arenas instead of being weight-decayed with a zero gradient
        named = [(k, p) for k, p in model.named_parameters() if p.requires_grad and not k.startswith("y_embedder.")]
        self.slots, total, self.block_ranges, self.head_range, self.tail_range = arena_layout(
            [(k, p.numel()) for k, p in named], len(model.blocks))
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        with torch.no_grad():
            for k, p in named:
                view = self.flat_p[self.slots[k]].view_as(p)
                view.copy_(p.detach().float())
                p.data = view
        half = model.t_embedder.frequency_embedding_size // 2          # the reference's own fp32 expression (DiT_models.py:52-54)
        self._freqs = torch.exp(-np.log(10000.0) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(dev)
        self.last_terms: Optional[Tensor] = None
        self.use_graph = os.environ.get("B200DIFF_TRAIN_GRAPH", "1") != "0"
        self.graph_warmup = 2
        self._graphs: dict = {}
        self._graph_seen: dict = {}
        self.graph_kernels = 0          # kernels in the last captured step graph
        self.replayed_kernels = 0       # kernels launched through graph replays so far (the C-side counter only sees eager launches)

    # -- helpers -------------------------------------------------------------------------------
    def _grad_view(self, name: str, shape) -> Tensor:
        return self.flat_g[self.slots[name]].view(shape)

    def _lin(self, prefix: str) -> _Lin:
        lin = _Lin()
        params = self._params
        wmat = params[prefix + ".weight"].detach()
        wmat = wmat.reshape(wmat.shape[0], -1)
        lin.o, lin.i = wmat.shape
        lin.w = ops.pack_conv_weight(wmat, self.dtype)
        lin.wt = ops.pack_conv_weight(wmat.t().contiguous(), self.dtype)
        lin.b = params[prefix + ".bias"].detach()
        lin.gw = self._grad_view(prefix + ".weight", (lin.o, lin.i))
        lin.gb = self._grad_view(prefix + ".bias", (lin.o,))
        return lin

    def _coef_rows(self, t: Tensor) -> Tensor:
        """Per-sample table rows for b200_q_sample / b200_diffusion_loss, gathered on the host from the float64 tables of
        the diffusion process and cast to fp32 exactly like _extract_into_tensor (gaussian_diffusion.py:1097-1109)."""
        g = self.diffusion
        ti = t.detach().cpu().numpy().astype(np.int64)
        rows = np.zeros((len(ti), 12), dtype=np.float64)
        rows[:, 0] = np.asarray(g.sqrt_alphas_cumprod)[ti]
        rows[:, 1] = np.asarray(g.sqrt_one_minus_alphas_cumprod)[ti]
        rows[:, 2] = np.asarray(g.posterior_mean_coef1)[ti]
        rows[:, 3] = np.asarray(g.posterior_mean_coef2)[ti]
        rows[:, 4] = np.asarray(g.posterior_log_variance_clipped)[ti]
        rows[:, 5] = np.log(np.asarray(g.betas))[ti]
        rows[:, 6] = (ti == 0)
        rows[:, 7] = (g.num_timesteps / 1000.0) if self.rescaled_mse else 1.0
        rows[:, 8] = np.asarray(g.sqrt_recip_alphas_cumprod)[ti]
        rows[:, 9] = np.asarray(g.sqrt_recipm1_alphas_cumprod)[ti]
        return torch.from_numpy(rows.astype(np.float32)).to(self.flat_p.device)

    def _model_t(self, t: Tensor) -> Tensor:
        """What _WrappedModel / _scale_timesteps hand to the network (respace.py:123-128)."""
        g = self.diffusion
        return torch.tensor([g._model_timestep(int(i)) for i in t.detach().cpu()], dtype=torch.float32,
                            device=self.flat_p.device)

    # -- the step ------------------------------------------------------------------------------
    @torch.no_grad()
    def training_step(self, x_start: Tensor, t: Tensor, cond: Optional[Sequence[Tensor]] = None,
                      noise: Optional[Tensor] = None, update: bool = True, use_graph: Optional[bool] = None) -> Dict[str, Tensor]:
        """One optimisation step on a batch: returns {'loss', 'mse', 'vb'} (batch means, like the trainer logs them).
        ``update=False`` stops after the gradient all-reduce (gradient parity tests read ``grads()``).

        ``use_graph`` (default ``self.use_graph``): after ``graph_warmup`` eager steps of a given batch shape the whole step —
        q_sample, forward, objective, backward, bucketed all-reduce, AdamW: ~2 000 launches — is captured ONCE as a CUDA
        graph and replayed; per step only the inputs, the per-sample table rows and three optimizer scalars are copied into
        static buffers.  Removes the launch gaps between the many small kernels (47 -> 42 ms per DiT-B/2 step)."""
        dev = self.flat_p.device
        x_start = x_start.to(dev).float().contiguous()
        n = x_start.shape[0]
        if n > 64:
            raise L.B200Error("DiTTrainer: at most 64 samples per GPU per step (conditioning-path kernels)")
        cond = [] if cond is None else [c.to(dev).float().contiguous() for c in cond]
        noise = None if noise is None else noise.to(dev).float().contiguous()
        coef, t_model = self._coef_rows(t), self._model_t(t)
        use_graph = self.use_graph if use_graph is None else use_graph
        # single-process only: a captured graph that holds NCCL collectives made torch.distributed's shutdown hang on the
        # 2-GPU box (the step itself replayed fine: 46.3 ms); data-parallel runs launch eagerly, where the all-reduce overlap
        # already hides the gaps
        use_graph = use_graph and not self._distributed()
        key = (tuple(x_start.shape), tuple(tuple(c.shape) for c in cond), noise is None)
        self._named = list(self.model.named_parameters())
        self._params = dict(self._named)
        if not (use_graph and update):
            return self._step(x_start, cond, noise, coef, t_model, update, None)
        st = self._graphs.get(key)
        if st is None or st["params"] != tuple(p.data_ptr() for _, p in self._named):
            seen = self._graph_seen.get(key, 0)
            if seen < self.graph_warmup:              # lazy one-time work (TMA descriptor cache, attribute sets, NCCL) runs eagerly
                self._graph_seen[key] = seen + 1
                return self._step(x_start, cond, noise, coef, t_model, True, None)
            st = self._capture(key, x_start, cond, noise, coef, t_model)
        st["x"].copy_(x_start)
        for b, c in zip(st["cond"], cond):
            b.copy_(c)
        if noise is not None:
            st["noise"].copy_(noise)
        st["coef"].copy_(coef)
        st["t_model"].copy_(t_model)
        self.step_count += 1
        st["hyper"].copy_(torch.tensor(T.adamw_hyper(self.lr, self.betas, self.step_count), dtype=torch.float32))
        st["graph"].replay()
        self.replayed_kernels += st["kernels"]
        self.model.invalidate()
        self.last_terms = st["terms"]
        return st["out"]

    def _capture(self, key, x_start, cond, noise, coef, t_model) -> dict:
        dev = self.flat_p.device
        st = {"x": x_start.clone(), "cond": [c.clone() for c in cond], "noise": None if noise is None else noise.clone(),
              "coef": coef.clone(), "t_model": t_model.clone(), "hyper": torch.zeros(3, device=dev),
              "params": tuple(p.data_ptr() for _, p in self._named)}
        if self._distributed():
            dist.barrier(self.group)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(g):
            st["out"] = self._step(st["x"], st["cond"], st["noise"], st["coef"], st["t_model"], True, st["hyper"])
            st["terms"] = self.last_terms
        st["graph"] = g
        st["kernels"] = L.launch_count() - n0          # libb200diff launches captured in (= replayed by) this graph
        self.graph_kernels = st["kernels"]
        self._graphs[key] = st
        return st

    def _step(self, x_start, cond, noise, coef, t_model, update: bool, hyper: Optional[Tensor]) -> Dict[str, Tensor]:
        m = self.model
        noise = torch.randn_like(x_start) if noise is None else noise
        x_t = T.q_sample(x_start, noise, coef)
        xin = torch.cat([x_t] + list(cond), dim=1) if cond else x_t
        self.flat_g.zero_()
        out, saved = self._forward(xin, t_model)
        learn_sigma = out.shape[1] == 2 * x_start.shape[1]
        terms, d_out = T.diffusion_loss(out, x_start, x_t, noise, coef, self.param, learn_sigma)
        self.last_terms = terms
        handles = self._backward(d_out, saved)
        for hnd in handles:
            hnd.wait()
        if update:
            world = dist.get_world_size(self.group) if self._distributed() else 1
            if hyper is None:
                self.step_count += 1
            T.adamw(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas, self.eps, self.weight_decay,
                    max(self.step_count, 1), 1.0 / world, hyper=hyper)
            m.invalidate()                               # packed sampling weights are stale now
        tm = terms.mean(0)
        return {"loss": tm[0] + tm[1], "mse": tm[0], "vb": tm[1]}

    def grads(self) -> Dict[str, Tensor]:
        return {k: self.flat_g[sl].view_as(p) for (k, p), sl in ((kp, self.slots[kp[0]]) for kp in self._named
                                                                   if kp[0] in self.slots)}

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce(self, rng: slice, handles: list) -> None:
        reduce_bucket(self.flat_g, rng, self.group, handles)

    # -- forward with saved activations --------------------------------------------------------
    def _forward(self, x: Tensor, t_model: Tensor):
        m, dt = self.model, self.dtype
        n, _, hh, ww = x.shape
        p, d = m.patch_size, m.hidden_size
        gh, gw = hh // p, ww // p
        sv: dict = {"n": n, "gh": gh, "gw": gw}
        te = m.t_embedder
        t_freq = ops.timestep_embedding(t_model.contiguous(), self._freqs, n, te.frequency_embedding_size, None)
        hid_pre = ops.linear_f32(t_freq, _f(te.mlp[0].weight), _f(te.mlp[0].bias))
        c = ops.linear_f32(hid_pre, _f(te.mlp[2].weight), _f(te.mlp[2].bias), in_silu=True)
        mod_w = torch.cat([_f(b.adaLN_modulation[1].weight) for b in m.blocks]
                          + [_f(m.final_layer.adaLN_modulation[1].weight)], 0).contiguous()
        mod_b = torch.cat([_f(b.adaLN_modulation[1].bias) for b in m.blocks]
                          + [_f(m.final_layer.adaLN_modulation[1].bias)], 0).contiguous()
        mod = ops.linear_f32(c, mod_w, mod_b, in_silu=True)
        sv.update(t_freq=t_freq, hid_pre=hid_pre, c=c, mod=mod, mod_w=mod_w)

        tok = ops.patchify(x.float(), p, dt)                                  # [N, T, kpad]
        wpe = m.x_embedder.proj.weight.detach()
        k = wpe[0].numel()
        kpad = tok.shape[-1]
        flat = torch.zeros((d, kpad), device=x.device)
        flat[:, :k] = wpe.reshape(d, k)
        pos = m.pos_embed.detach().to(dt).expand(n, -1, -1).contiguous()
        h = linear_tokens(tok, ops.pack_conv_weight(flat, dt), bias=_f(m.x_embedder.proj.bias), residual=pos)
        sv.update(tok=tok, k=k)
        heads, dh = m.num_heads, d // m.num_heads
        blocks = []
        for i, blk in enumerate(m.blocks):
            pre = f"blocks.{i}."
            lq, lp = self._lin(pre + "attn.qkv"), self._lin(pre + "attn.proj")
            l1, l2 = self._lin(pre + "mlp.fc1"), self._lin(pre + "mlp.fc2")
            mm = mod[:, i * 6 * d:(i + 1) * 6 * d]
            sl = lambda j: mm[:, j * d:(j + 1) * d]
            a1 = ops.layer_norm(h, None, None, 1e-6, shift=sl(0), scale=sl(1))
            qkv = linear_tokens(a1, lq.w, bias=lq.b)
            lse = torch.empty((n, heads, qkv.shape[1]), device=x.device, dtype=torch.float32)
            att = ops.attention(qkv, qkv, qkv, heads, dh, blk.attn.scale, q_off=0, k_off=d, v_off=2 * d, lse=lse)
            f1 = linear_tokens(att, lp.w, bias=lp.b)
            h1 = T.gated_residual(h, f1, sl(2))
            a2 = ops.layer_norm(h1, None, None, 1e-6, shift=sl(3), scale=sl(4))
            pre_act = linear_tokens(a2, l1.w, bias=l1.b)
            u = T.gelu_fwd(pre_act)
            f2 = linear_tokens(u, l2.w, bias=l2.b)
            h2 = T.gated_residual(h1, f2, sl(5))
            blocks.append(dict(h0=h, a1=a1, qkv=qkv, lse=lse, att=att, f1=f1, h1=h1, a2=a2, pre=pre_act, u=u, f2=f2,
                               lq=lq, lp=lp, l1=l1, l2=l2))
            h = h2
        fm = mod[:, len(m.blocks) * 6 * d:]
        af = ops.layer_norm(h, None, None, 1e-6, shift=fm[:, :d], scale=fm[:, d:])
        wf, bf = m.final_layer.linear.weight.detach(), m.final_layer.linear.bias.detach()
        cols = wf.shape[0]
        cp = (cols + 15) // 16 * 16
        wfp = torch.zeros((cp, d), device=x.device)
        wfp[:cols] = wf
        bfp = torch.zeros(cp, device=x.device)
        bfp[:cols] = bf
        outt = linear_tokens(af, ops.pack_conv_weight(wfp, dt), bias=bfp, out_f32=True)
        out = ops.unpatchify(outt, m.out_channels, gh, gw, p)
        sv.update(blocks=blocks, h_last=h, af=af, wfp=wfp, cols=cols, cp=cp)
        return out, sv

    # -- backward ------------------------------------------------------------------------------
    def _lin_bwd(self, lin: _Lin, dy: Tensor, x: Tensor, need_dx: bool = True) -> Optional[Tensor]:
        """dy [N, L, O], x [N, L, I] -> accumulates dW, db into the gradient arena, returns dx [N, L, I]."""
        n, l, o = dy.shape
        dy2, x2 = dy.view(n * l, o), x.view(n * l, -1)
        T.linear_wgrad(dy2, x2, lin.gb, out=lin.gw)
        return linear_tokens(dy, lin.wt) if need_dx else None

    def _backward(self, d_out: Tensor, sv: dict) -> list:
        m, dt = self.model, self.dtype
        n, d, p = sv["n"], m.hidden_size, m.patch_size
        depth = len(m.blocks)
        heads, dh = m.num_heads, d // m.num_heads
        mod = sv["mod"]
        dmod = torch.zeros_like(mod)
        handles: list = []
        # final layer
        cp, cols = sv["cp"], sv["cols"]
        dtok = T.unpatchify_bwd(d_out, p, cp, dt)                                   # [N, T, cp]
        t_all = dtok.shape[0] * dtok.shape[1]
        gbf = torch.zeros(cp, device=dtok.device)
        gwf = T.linear_wgrad(dtok.view(t_all, cp), sv["af"].view(t_all, d), gbf)    # [cp, d]
        self._grad_view("final_layer.linear.weight", (cols, d)).copy_(gwf[:cols])
        self._grad_view("final_layer.linear.bias", (cols,)).copy_(gbf[:cols])
        daf = linear_tokens(dtok, ops.pack_conv_weight(sv["wfp"].t().contiguous(), dt))     # [N, T, d]
        fo = depth * 6 * d
        dh_ = T.ln_mod_bwd(sv["h_last"], daf, mod[:, fo + d:fo + 2 * d], None, dmod[:, fo:fo + d], dmod[:, fo + d:fo + 2 * d])
        # blocks, last to first
        for i in reversed(range(depth)):
            b = sv["blocks"][i]
            o = i * 6 * d
            ms = lambda j: mod[:, o + j * d:o + (j + 1) * d]
            dm = lambda j: dmod[:, o + j * d:o + (j + 1) * d]
            df2 = T.gate_bwd(dh_, b["f2"], ms(5), dm(5))
            du = self._lin_bwd(b["l2"], df2, b["u"])
            dpre = T.gelu_bwd(b["pre"], du)
            da2 = self._lin_bwd(b["l1"], dpre, b["a2"])
            dh1 = T.ln_mod_bwd(b["h1"], da2, ms(4), dh_, dm(3), dm(4))
            df1 = T.gate_bwd(dh1, b["f1"], ms(2), dm(2))
            datt = self._lin_bwd(b["lp"], df1, b["att"])
            dqkv = T.attention_bwd(b["qkv"], b["att"], datt, b["lse"], heads, dh, m.blocks[i].attn.scale)
            da1 = self._lin_bwd(b["lq"], dqkv, b["a1"])
            dh_ = T.ln_mod_bwd(b["h0"], da1, ms(1), dh1, dm(0), dm(1))
            # this block's adaLN projection: dW = dmod_i^T silu(c), db = sum_n dmod_i  (fp32, tiny)
            pre = f"blocks.{i}.adaLN_modulation.1"
            T.linear_f32_wgrad(dmod[:, o:o + 6 * d], sv["c"], self._grad_view(pre + ".weight", (6 * d, d)),
                               self._grad_view(pre + ".bias", (6 * d,)), in_silu=True)
            self._reduce(self.block_ranges[i], handles)        # bucket complete: overlap its all-reduce with the rest
        # patch embedding: h = tok W^T + b + pos
        t_all = dh_.shape[0] * dh_.shape[1]
        gpe = T.linear_wgrad(dh_.view(t_all, d), sv["tok"].view(t_all, -1),
                             self._grad_view("x_embedder.proj.bias", (d,)))        # [d, kpad]
        k = sv["k"]
        self._grad_view("x_embedder.proj.weight", (d, k)).copy_(gpe[:, :k])
        # conditioning path: final adaLN projection, then mod -> c -> TimestepEmbedder MLP
        pre = "final_layer.adaLN_modulation.1"
        T.linear_f32_wgrad(dmod[:, fo:fo + 2 * d], sv["c"], self._grad_view(pre + ".weight", (2 * d, d)),
                           self._grad_view(pre + ".bias", (2 * d,)), in_silu=True)
        dc = T.linear_f32_dgrad(dmod, sv["mod_w"], x_pre=sv["c"])
        te = m.t_embedder
        T.linear_f32_wgrad(dc, sv["hid_pre"], self._grad_view("t_embedder.mlp.2.weight", tuple(te.mlp[2].weight.shape)),
                           self._grad_view("t_embedder.mlp.2.bias", tuple(te.mlp[2].bias.shape)), in_silu=True)
        dhid = T.linear_f32_dgrad(dc, _f(te.mlp[2].weight), x_pre=sv["hid_pre"])
        T.linear_f32_wgrad(dhid, sv["t_freq"], self._grad_view("t_embedder.mlp.0.weight", tuple(te.mlp[0].weight.shape)),
                           self._grad_view("t_embedder.mlp.0.bias", tuple(te.mlp[0].bias.shape)))
        self._reduce(self.head_range, handles)
        self._reduce(self.tail_range, handles)
        return handles
