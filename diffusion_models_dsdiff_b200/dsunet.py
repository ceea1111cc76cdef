"""Drop-in ``DSUnetModel`` — the DS-Diff four-stream backbone (UNet_DS_Diff/model.py:171-756).

Same constructor keywords, ``forward(x, timesteps)`` -> ``(out, feature_dict)`` and state_dict keys as
the reference: four encoders with separate weights (``input_blocks``, ``input_blocks_a``, ``_al``,
``_l``; :522-524) fed by the four input channels (two-channel input zero-fills the last two streams,
:654-658), the middle block on the noise stream only (:688), four ``FeatureDisentangle`` heads applied
12 times (:695-719), SE_Attention + 3x3 projections of the per-family means (:722-725), a 1x1
``all_proj`` over SiLU of the 5-way concat (:734-738) and a decoder whose skip is the mean of the four
encoders' skips (:745).

All arithmetic runs through the same sm_100a kernels as ``UNetModel`` (see unet.py): the skip mean is
one fused 4-input kernel whose output then enters the concat-aware GroupNorm / split-K convolution, so
``cat([h, skip])`` is never materialised.
"""
from __future__ import annotations

import copy
import os
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .unet import (Act, Ctx, ResBlock, _EngineModule, _f, _gn, build_unet_trunk, run_conv, run_gn, run_in_conv,
                   run_out_head, run_resblock_conv1_raw, run_sequential)

Tensor = torch.Tensor


class SE_Attention(nn.Module):
    """Parameters of Disc_diff/guided_diffusion/unet.py:82-109 (bias-free squeeze/excite MLP)."""

    def __init__(self, channel=512, reduction=16):
        super().__init__()
        self.se = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())


class FeatureDisentangle(nn.Module):
    """Parameters of UNet_DS_Diff/model.py:152-168."""

    def __init__(self, in_channels, half_conv_ch):
        super().__init__()
        self.conv_1 = nn.Sequential(_gn(in_channels), nn.SiLU(), nn.Conv2d(in_channels, in_channels, 3, 1, 1))
        self.conv_2 = nn.Sequential(_gn(in_channels), nn.SiLU(), nn.Conv2d(in_channels, half_conv_ch, 1, 1, 0))


class DSUnetModel(_EngineModule):
    _returns_features = True

    def __init__(self, *args, **kwargs):
        super().__init__()
        build_unet_trunk(self, *args, **kwargs)
        self.input_blocks_a = nn.ModuleList([copy.deepcopy(m) for m in self.input_blocks])
        self.input_blocks_al = nn.ModuleList([copy.deepcopy(m) for m in self.input_blocks])
        self.input_blocks_l = nn.ModuleList([copy.deepcopy(m) for m in self.input_blocks])
        input_ch = int(self.channel_mult[0] * self.model_channels)
        conv_ch = input_ch * self.channel_mult[-1]
        half = int(conv_ch / 2)
        self.conv_style = FeatureDisentangle(conv_ch, half)
        self.conv_content = FeatureDisentangle(conv_ch, half)
        self.conv_anatomy = FeatureDisentangle(conv_ch, half)
        self.conv_lesion = FeatureDisentangle(conv_ch, half)
        mk = lambda: nn.Sequential(SE_Attention(half, reduction=8), nn.Conv2d(half, half, 3, 1, 1))
        self.style_proj, self.share_content_proj = mk(), mk()
        self.anatomy_proj, self.lesion_proj = mk(), mk()
        self.all_proj = nn.Sequential(nn.SiLU(), nn.Conv2d(half * 6, conv_ch, 1, 1, 0))
        self._init_engine()
        self.encoder_streams = os.environ.get("B200DIFF_ENCODER_STREAMS", "1") != "0"
        self._streams = {}
        # loop-invariant prefix of the conditioning encoders (see prepare_conditioning)
        self.hoist_conditioning = os.environ.get("B200DIFF_HOIST_COND", "1") != "0"
        self._cond_caches = {}
        self._cond_key = None

    def invalidate(self):
        super().invalidate()
        self._cond_key = None

    @torch.no_grad()
    def prepare_conditioning(self, cond) -> bool:
        """Called by the sampling loops once per run with the conditioning tensors they will concatenate to x_t at EVERY
        step (DiffusionWrapper 'concat', ddpm.py:1331-1333).  The three conditioning encoders see the same input at every
        step, and a ResBlock only meets the timestep after its first convolution (openaimodel.py:271-282), so
        ``input_blocks_X[0]`` (first conv), and GN -> SiLU -> conv1 of ``input_blocks_X[1]`` are loop invariants: they
        are evaluated here once per run — not 50 (or 2000) times — into persistent buffers that the per-step graph reads.
        Each step then adds the bias + timestep term inside the next GroupNorm (statistics corrected analytically).
        Exactly the reference's arithmetic up to the rounding point of the cached 16-bit tensor.  Returns True if the
        cache is armed (``forward(..., cond_cached=True)`` will use it)."""
        self._cond_key = None
        if not self.hoist_conditioning or cond is None:
            return False
        cond = list(cond) if isinstance(cond, (list, tuple)) else [cond]
        c = torch.cat([t.float() for t in cond], dim=1) if len(cond) > 1 else cond[0].float()
        if c.dim() != 4 or c.shape[1] not in (1, 3) or not c.is_cuda:
            return False
        if not isinstance(self.input_blocks[1][0], ResBlock):
            return False
        pw = self.packed()
        n = c.shape[0]
        streams = [c[:, 0:1]] + ([c[:, 1:2], c[:, 2:3]] if c.shape[1] == 3 else [torch.zeros_like(c[:, 0:1])] * 2)
        key = (n, c.shape[2], c.shape[3], self.act_dtype, c.device.index)
        arena = ops.StatsArena()
        arena.begin(c.device)
        cx = Ctx(pw, None, arena)
        fresh = []
        for blocks, xin in zip((self.input_blocks_a, self.input_blocks_al, self.input_blocks_l), streams):
            h0 = run_in_conv(cx, blocks[0][0], xin.contiguous(), self.act_dtype)
            if h0.cs is None:
                h0.cs = ops.channel_stats(h0.t, cx.stats_for(n, h0.t.shape[-1]))
            c1 = run_resblock_conv1_raw(cx, blocks[1][0], h0)
            fresh.append((h0, c1))
        slot = self._cond_caches.get(key)
        if slot is None:        # persistent buffers (captured graphs keep pointing at them); statistics outside the arena
            slot = [(Act(h0.t.clone(), h0.cs.clone()), Act(c1.t.clone(), c1.cs.clone())) for h0, c1 in fresh]
            self._cond_caches[key] = slot
        else:
            for (ph0, pc1), (h0, c1) in zip(slot, fresh):
                ph0.t.copy_(h0.t); ph0.cs.copy_(h0.cs); pc1.t.copy_(c1.t); pc1.cs.copy_(c1.cs)
        self._cond_key = key
        return True

    def _side_streams(self, device):
        key = (device.type, device.index)
        if key not in self._streams:
            self._streams[key] = [torch.cuda.Stream(device=device) for _ in range(4)]
        return self._streams[key]

    # FeatureDisentangle.forward (:165-168): conv_2(conv_1(x) + x)
    def _fd(self, cx: Ctx, m: FeatureDisentangle, x: Act) -> Tensor:
        pw = cx.pw
        g1, c1, g2, c2 = m.conv_1[0], m.conv_1[2], m.conv_2[0], m.conv_2[2]
        a = run_gn(cx, g1, x)
        out = run_conv(cx, [(a, pw.conv[id(c1)], 0, 3, 1)], bias=_f(c1.bias), residual=x.t)
        a2 = run_gn(cx, g2, out)
        return ops.conv2d([(a2, pw.conv[id(c2)], 0, 1, 0)], bias=_f(c2.bias))

    # mean of a feature family -> SE_Attention -> 3x3 conv (:722-725)
    def _proj(self, cx: Ctx, seq: nn.Sequential, feats) -> Tensor:
        mean = ops.add4_scale(*feats, *([None] * (4 - len(feats))), scale=1.0 / len(feats))
        se, conv = seq[0].se, seq[1]
        gated = ops.se_attention(mean, _f(se[0].weight), _f(se[2].weight))
        return ops.conv2d([(gated, cx.pw.conv[id(conv)], 0, 3, 1)], bias=_f(conv.bias))

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, step_ptr: Optional[Tensor] = None,
                return_features: bool = True, cond_cached: bool = False, **kwargs):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if context is not None:
            raise NotImplementedError("B200 engine: cross-attention context is not built yet")
        pw = self.packed()
        n = x.shape[0]
        emb_all = self._emb_all(pw, timesteps, n, step_ptr)
        self._arena.begin(x.device)
        cx = Ctx(pw, emb_all, self._arena)
        x = x.float()
        if x.shape[1] == 2:       # :654-658
            zero = torch.zeros_like(x[:, 0:1])
            inputs = [x[:, 0:1], x[:, 1:2], zero, zero]
        else:                     # :660-663
            inputs = [x[:, 0:1], x[:, 1:2], x[:, 2:3], x[:, 3:4]]
        cached = None
        if cond_cached and self._cond_key is not None and self._cond_key[:3] == (n, x.shape[2], x.shape[3]):
            cached = self._cond_caches[self._cond_key]      # armed by prepare_conditioning() for THIS run's cond

        def encoder(blocks, xin, pre=None):
            if pre is not None:           # (first conv, conv1 of the first ResBlock) from the per-run cache
                h, c1 = pre
                hs = [h]
                h = run_sequential(cx, blocks[1], h, None, conv1_raw=c1)
                hs.append(h)
                rest = list(blocks)[2:]
            else:
                h = run_in_conv(cx, blocks[0][0], xin.contiguous(), self.act_dtype)
                hs = [h]
                rest = list(blocks)[1:]
            for blk in rest:
                h = run_sequential(cx, blk, h, None)
                hs.append(h)
            return hs

        enc_blocks = (self.input_blocks, self.input_blocks_a, self.input_blocks_al, self.input_blocks_l)
        # The four encoders are independent (separate weights, same emb): the three conditioning encoders run on side
        # streams, forked after the timestep path and joined before their first consumer, so their latency-bound
        # 8x8 .. 32x32 layers overlap with the noise stream's instead of leaving most SMs idle.  Inside a captured
        # CUDA graph this becomes four parallel branches.  (Per-launch event timing keeps everything on one stream.)
        fork = self.encoder_streams and ops._timer is None
        if fork:
            main = torch.cuda.current_stream(x.device)
            side = self._side_streams(x.device)
            ev0 = torch.cuda.Event()
            ev0.record(main)
            skips = [None] * 4
            joins = []
            for i in (1, 2, 3):
                with torch.cuda.stream(side[i - 1]):
                    side[i - 1].wait_event(ev0)
                    skips[i] = encoder(enc_blocks[i], inputs[i], None if cached is None else cached[i - 1])
                    ev = torch.cuda.Event()
                    ev.record(side[i - 1])
                    joins.append(ev)
            skips[0] = encoder(enc_blocks[0], inputs[0])
        else:
            skips = [encoder(b, xin, None if (cached is None or i == 0) else cached[i - 1])
                     for i, (b, xin) in enumerate(zip(enc_blocks, inputs))]
        last = [hs[-1] for hs in skips]

        def skip_mean(level_acts):
            s0 = level_acts[0].t
            st = cx.stats_for(s0.shape[0], s0.shape[-1])
            return Act(ops.add4_scale_stats(s0, level_acts[1].t, level_acts[2].t, level_acts[3].t, 0.25, st), st)  # :745

        skip_means = None
        if fork and os.environ.get("B200DIFF_SKIP_MEAN_STREAM", "1") != "0":
            # the decoder's skip inputs (mean of the four encoders' skips) only need the encoders: all of them are computed
            # on a fourth stream right away, so these HBM-bound passes overlap the latency-bound 8x8 middle block and feature
            # heads instead of sitting between the decoder's convolutions
            ev_enc = torch.cuda.Event()
            ev_enc.record(main)                      # the noise encoder (and everything before it) on the main stream
            with torch.cuda.stream(side[3]):
                side[3].wait_event(ev_enc)
                for ev in joins:
                    side[3].wait_event(ev)
                skip_means = [skip_mean([skips[e][lvl] for e in range(4)]) for lvl in range(len(skips[0]))]
                ev_means = torch.cuda.Event()
                ev_means.record(side[3])
        h_n = run_sequential(cx, self.middle_block, last[0], None)     # :688
        h_a, h_al, h_l = last[1], last[2], last[3]
        fd = lambda m, v: self._fd(cx, m, v)
        if fork:
            # the four feature families (3 + 3 + 2 + 2 FeatureDisentangle heads on 8x8 maps, then mean -> SE -> conv) are
            # independent chains of latency-bound kernels: one family per stream, joined before the 5-way concat
            for ev in joins:
                main.wait_event(ev)
            ev1 = torch.cuda.Event()
            ev1.record(main)
            fam = {}

            def family(name, head, proj, srcs):
                feats = [fd(head, v) for v in srcs]
                fam[name] = (feats, self._proj(cx, proj, feats))

            work = [("content", self.conv_content, self.share_content_proj, (h_a, h_al, h_l)),
                    ("anatomy", self.conv_anatomy, self.anatomy_proj, (h_a, h_al)),
                    ("lesion", self.conv_lesion, self.lesion_proj, (h_al, h_l))]
            joins2 = []
            for st, args in zip(side, work):
                with torch.cuda.stream(st):
                    st.wait_event(ev1)
                    family(*args)
                    ev = torch.cuda.Event()
                    ev.record(st)
                    joins2.append(ev)
            family("style", self.conv_style, self.style_proj, (h_a, h_al, h_l))
            # style / content heads of the NOISE stream only feed the returned feature dict (model.py:695-700, 751-756):
            # every sampler discards it, so they are skipped when the caller does not ask for features
            h_n_style, h_n_content = (fd(self.conv_style, h_n), fd(self.conv_content, h_n)) if return_features else (None, None)
            for ev in joins2:
                main.wait_event(ev)
            (style, h_style), (content, h_share) = fam["style"], fam["content"]
            (anatomy, h_anat), (lesion, h_les) = fam["anatomy"], fam["lesion"]
        else:
            # style / content heads of the NOISE stream only feed the returned feature dict (model.py:695-700, 751-756):
            # every sampler discards it, so they are skipped when the caller does not ask for features
            h_n_style, h_n_content = (fd(self.conv_style, h_n), fd(self.conv_content, h_n)) if return_features else (None, None)
            style = [fd(self.conv_style, v) for v in (h_a, h_al, h_l)]
            content = [fd(self.conv_content, v) for v in (h_a, h_al, h_l)]
            anatomy = [fd(self.conv_anatomy, v) for v in (h_a, h_al)]
            lesion = [fd(self.conv_lesion, v) for v in (h_al, h_l)]
            h_style = self._proj(cx, self.style_proj, style)
            h_share = self._proj(cx, self.share_content_proj, content)
            h_anat = self._proj(cx, self.anatomy_proj, anatomy)
            h_les = self._proj(cx, self.lesion_proj, lesion)
        cat = torch.cat([h_n.t, h_share, h_style, h_anat, h_les], dim=-1)          # :734 (8x8 maps: tiny)
        ap = self.all_proj[1]
        h = run_conv(cx, [(ops.add4_scale(cat, silu=True), pw.conv[id(ap)], 0, 1, 0)], bias=_f(ap.bias))   # :738
        if skip_means is not None:
            main.wait_event(ev_means)
        for blk in self.output_blocks:
            if skip_means is not None:
                skip = skip_means.pop()
            else:
                skip = skip_mean([skips[e].pop() for e in range(4)])
            h = run_sequential(cx, blk, h, skip)
        out = run_out_head(cx, self.out, h).to(x.dtype)
        self._range_check(x.device, step_ptr)
        if not return_features:
            return out, None
        to = ops.nhwc_to_nchw
        feats = {"style": [to(v) for v in style], "content": [to(v) for v in content],
                 "anatomy": [to(v) for v in anatomy], "lesion": [to(v) for v in lesion],
                 "n_style_content": [to(h_style), to(h_n_style), to(h_share), to(h_n_content)]}
        return out, feats
