/*
 * b200diff.h — C ABI of libb200diff.so (sm_100a kernels for the DS-Diff sampling hot path).
 *
 * The reference (larrybb626/diffusion_models_dsdiff) is pure Python/PyTorch and has NO FFI of
 * its own: the seam a maintainer would bind is the set of torch ops its
 * hot path dispatches.  Each entry point below names the reference call site(s) whose arithmetic
 * it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers + sizes only; no torch / C++ types.  All data pointers are DEVICE pointers
 *     unless the name ends in _host.  `stream` is a cudaStream_t passed as void*.
 *   - never allocates device memory, never synchronises, never falls back to the CPU.
 *   - returns 0 on success; non-zero = error, text via b200_last_error() (thread-local).
 *   - activations inside the network are NHWC 16-bit ("act"): B200_F16 (IEEE half, default —
 *     same 10-bit mantissa as the TF32 the reference runs with) or B200_BF16.  Tensor-core
 *     accumulation, GroupNorm statistics, softmax, time embedding and the sampler state are fp32.
 *   - the API edge (what the Python drop-in classes expose) is NCHW fp32 like the reference.
 */
#ifndef B200DIFF_H_
#define B200DIFF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 16
#define B200_F16 0
#define B200_BF16 1

/* ---- library ---------------------------------------------------------------------------- */
int b200_abi_version(void);
const char *b200_last_error(void);
/* one-time per-process init on the CURRENT device (driver entry points, smem attributes). */
int b200_init(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t b200_launch_count(void);

/* ---- layout conversion at the API edge ----------------------------------------------------
 * replaces nothing in the reference (it is NCHW fp32 throughout); needed because the kernels
 * run NHWC 16-bit.  src NCHW fp32 [N,C,H,W] -> dst NHWC act [N,H,W,Cpad] (channels >= C are 0). */
int b200_nchw_f32_to_nhwc(const float *src, void *dst, int N, int C, int H, int W, int Cpad,
                          int act_dtype, void *stream);
int b200_nhwc_to_nchw_f32(const void *src, float *dst, int N, int C, int H, int W, int Csrc,
                          int act_dtype, void *stream);

/* ---- weight repack (once per load_state_dict) ----------------------------------------------
 * OIHW fp32 [Cout,Cin,KH,KW] (state_dict layout, ldm/modules/diffusionmodules/util.py:229-239
 * conv_nd) -> [KH*KW][Cout][CinPad] act dtype, K-major rows for the tcgen05 B operand.       */
int b200_pack_conv_weight(const float *w_oihw, void *dst, int Cout, int Cin, int KH, int KW,
                          int CinPad, int act_dtype, void *stream);

/* ---- implicit-GEMM convolution on tcgen05/TMEM, operands by TMA ------------------------------
 * replaces F.conv2d as called by ResBlock in_layers[2]/out_layers[3]/skip_connection
 * (ldm/modules/diffusionmodules/openaimodel.py:208,234,245), Downsample.op (:155-157),
 * Upsample.conv (:118-120), AttentionBlock qkv/proj_out Conv1d (:452,460), FeatureDisentangle and
 * the *_proj/all_proj convs of DSUnetModel (UNet_DS_Diff/model.py:158,163,569-601), and every
 * nn.Linear applied to a [B,L,C] token tensor (ldm/modules/attention.py, DiT_models.py).
 *
 *   out[n,ho,wo,co] = sum_seg sum_{r,s,ci} src_seg[n, ho*stride+r-pad, wo*stride+s-pad, ci]
 *                                          * w_seg[r*k+s][co][coff_seg+ci]
 *                     + bias[co] + rowbias[n*rowbias_stride + co] + res_scale*residual[n,ho,wo,co]
 *
 * Up to 3 K-segments accumulate into the same TMEM tile: this is how concat inputs
 * (th.cat([h, hs.pop()],1), openaimodel.py:952) and the fused ResBlock 1x1 skip are expressed
 * without materialising the concatenation.  Channel counts must be multiples of 8.            */
typedef struct {
  const void *src;     /* NHWC act [N,Hin,Win,cin]                                   */
  const void *weight;  /* packed [k*k][Cout][wcin] act (b200_pack_conv_weight)       */
  int cin;             /* channels of src                                            */
  int wcin;            /* Cin pitch of the packed weight (>= coff+cin)               */
  int coff;            /* first weight input-channel used by this segment            */
  int ksize;           /* 1 or 3 (2: the 2x2 sub-pixel kernels of a folded Upsample) */
  int pad;             /* 0 or 1: zero padding (top / left)                          */
  int pad_w1;          /* 0: left padding = pad; k + 1: left padding = k (asymmetric 2x2 sub-pixel kernels) */
} b200_conv_seg;

typedef struct {
  int N, Hin, Win;     /* input spatial size (all segments share it)                 */
  int Hout, Wout;      /* output spatial size                                        */
  int Cout;
  int stride;          /* 1 or 2                                                     */
  int nseg;            /* 1..3                                                       */
  b200_conv_seg seg[3];
  const float *bias;       /* [Cout] fp32 or NULL                                    */
  const float *rowbias;    /* per-(n,co) fp32 add (timestep embedding, :282) or NULL */
  int rowbias_stride;      /* elements between consecutive n                         */
  const void *residual;    /* NHWC act [N,Hout,Wout,Cout] or NULL                    */
  float res_scale;
  void *out;               /* NHWC act [N,Hout,Wout,Cout]                            */
  double *gn_sums;         /* optional [N][Cout][2] fp64 (sum, sum of squares per image and channel),
                              ACCUMULATED from the fp32 epilogue values for a following GroupNorm
                              (zero it first), or NULL                                */
  int act_dtype;
  float *out_nchw_f32;     /* optional: write channels [0, cout_real) as NCHW fp32 [N,cout_real,Hout,Wout] INSTEAD of
                              `out` — the network's last conv (self.out[2], openaimodel.py:898-902) whose 1-2 output
                              channels are zero-padded to a 16-wide N tile; NULL otherwise                 */
  int cout_real;
  const float *rowscale;   /* optional per-(n,co) fp32 factor applied to (acc + bias + rowbias) BEFORE the residual add:
                              the adaLN-Zero gate of DiTBlock, x + gate * f(x) (UNet_DS_Diff/DiT_models.py:118-122) */
  int rowscale_stride;
  int epi_act;             /* activation on (acc + bias + rowbias): 0 none, 1 GELU(tanh) (timm Mlp in DiT, :111-112) */
  float *out_nhwc_f32;     /* optional: write the result as NHWC fp32 [N,Hout,Wout,Cout] INSTEAD of `out` (attention
                              logits q k^T of the KL-autoencoder AttnBlock, ldm/modules/diffusionmodules/model.py:186-189,
                              which must stay fp32 until the softmax); NULL otherwise                       */
  void *workspace;          /* split-K scratch (fp32 partial tiles), >= b200_conv_ksplit(d) * N*Hout*Wout*Cout*4 bytes,
                              16-byte aligned; only read / written by this call; may be NULL when the factor is 1      */
  size_t workspace_bytes;
  unsigned int *range_events; /* optional device counter (caller-owned, zeroed by the caller): incremented once per output
                              tile in which an fp16 result had to be SATURATED to +-65504 (|value| >= 65520, or inf).  The
                              reference computes in fp32 and cannot overflow here; a caller that sees a non-zero count must
                              treat the results as invalid (the Python layer raises and suggests bf16 activations).  NULL =
                              no counting.  Never written for B200_BF16 (fp32 exponent range).                          */
  int wsets;                /* 0 or 1: one weight set.  G > 1: "grouped along N" — the N images are G equal groups and group
                              g = n / (N / G) uses its own weights, bias and timestep rows: seg[i].weight is the stack
                              [G][k*k][Cout][wcin], bias [G][Cout], rowbias / rowscale advance by *_gstride floats per group
                              (their per-image stride then counts images WITHIN the group).  One launch for the G same-shape
                              encoders of DSUnetModel (UNet_DS_Diff/model.py:674-686) instead of G latency-bound ones.      */
  int rowbias_gstride, rowscale_gstride;
  int out_sy, out_sx, out_oy, out_ox;
                            /* 0 / 1: dense output.  out_sy, out_sx > 1: output pixel (h, w) of this launch is stored at
                              (h*out_sy + out_oy, w*out_sx + out_ox) of an NHWC tensor [N, Hout*out_sy, Wout*out_sx, Cout] —
                              one parity class of `Upsample` (nearest x2, openaimodel.py:113-117) followed by its 3x3 conv
                              (:118-120), which collapses to four 2x2 convolutions of the LOW-resolution input (one per output
                              parity, weights = sums of the 3x3 taps that read the same input pixel): the 4x-sized intermediate
                              is never written and 16 instead of 36 multiply-adds per output pixel and channel pair remain.
                              gn_sums of the four launches accumulate into the same buffer.                              */
  int force_ksplit;         /* 0: split-K factor chosen by b200_conv_ksplit (geometry only).  S > 1: run exactly S K-splits; with
                              out_nhwc_f32 set the S fp32 partial planes (workspace) are summed in fixed order into out_nhwc_f32
                              with no epilogue terms (bias / residual must be NULL) — the weight-gradient GEMM of the training
                              step, dW[Cout, Cin] = dY^T X with K = all tokens of the batch (trainers' loss.backward()).        */
  void *out_lo;             /* optional second 16-bit NHWC plane of `out`'s shape: receives result - (stored 16-bit result), so
                              that out + out_lo carries the fp32 value the reference keeps in its fp32 activation tensors
                              (ResBlock / first-conv outputs of the 256x256 level of DSUnetModel, whose 16-bit rounding is what
                              DESIGN.md section 4 measures).  Operand consumers (TMA) read `out` alone — it is exactly the 16-bit
                              rounding they would apply anyway; GroupNorm / residual consumers read both planes.  Needs a 16-bit
                              NHWC `out` and a launch without split-K (b200_conv_ksplit == 1).                              */
  const void *residual_lo;  /* optional low plane of `residual` (same shape / type): the epilogue adds residual + residual_lo */
} b200_conv_desc;
int b200_conv2d_igemm(const b200_conv_desc *d, void *stream);
/* split-K factor b200_conv2d_igemm will use for this descriptor (1 = none).  It depends on the layer geometry only
 * (Hout, Wout, channels, kernel sizes; never on N), so one image's result is the same in any batch; the caller sizes
 * `workspace` from it (the library never allocates).                                                                 */
int b200_conv_ksplit(const b200_conv_desc *d);

/* first conv of the network: tiny Cin (<= 8) read straight from NCHW fp32, writes NHWC act.
 * replaces input_blocks[0] conv (openaimodel.py:717) after DiffusionWrapper's concat
 * (ldm/models/diffusion/ddpm.py:1331-1333).  w is OIHW fp32.  sums (may be NULL): zeroed fp64 [N, Cout, 2], receives
 * the per-(image, channel) sum / sum of squares of the stored output for the following GroupNorm.     */
int b200_conv3x3_in(const float *x_nchw, const float *w_oihw, const float *bias, void *out_nhwc,
                    int N, int Cin, int H, int W, int Cout, double *sums, void *out_lo, int act_dtype, void *stream);
/* out_lo (may be NULL): low plane of the output, see b200_conv_desc.out_lo.                            */
/* last conv: Cout <= 4, NHWC act -> NCHW fp32.  replaces self.out[2] (openaimodel.py:898-902). */
int b200_conv3x3_out(const void *x_nhwc, const float *w_oihw, const float *bias, float *out_nchw,
                     int N, int Cin, int H, int W, int Cout, int act_dtype, void *stream);

/* ---- GroupNorm(+FiLM)(+SiLU), fused single apply pass ----------------------------------------
 * replaces GroupNorm32 -> SiLU (util.py:209-226, openaimodel.py:205-209,229-236,898-902), the
 * FiLM form GN(h)*(1+scale)+shift (openaimodel.py:276-280) and Normalize (eps 1e-6,
 * ldm/modules/attention.py:88-89).  Input may be the channel-concatenation of two tensors.
 * stats: [N][groups][2] fp64 (sum, sumsq), zero before b200_groupnorm_stats.                   */
int b200_groupnorm_stats(const void *x0, int C0, const void *x1, int C1, int N, int HW,
                         int groups, double *stats, int act_dtype, void *stream);
int b200_groupnorm_apply(const void *x0, int C0, const void *x1, int C1, int N, int HW,
                         int groups, float eps, const double *stats, const float *gamma,
                         const float *beta, const float *film, int film_stride, int silu,
                         void *out, int act_dtype, void *stream);

/* Producer-side statistics: sums[N][C][2] fp64 = per-(image, channel) sum and sum of squares, accumulated (zero first).
 * b200_conv2d_igemm fills the same layout from its epilogue (gn_sums); b200_groupnorm_apply_cs reduces the channel
 * sums of the two concatenated sources to group statistics itself, so a GroupNorm is ONE read + ONE write pass.
 * preshift (may be NULL): fp32 [N or 1][C0] row pitch preshift_stride (0 = one row for the batch): the tensor normalised
 * is x0 + preshift[n, c]; the sums of x0 are corrected analytically.  Used for `h + emb_out` of ResBlock._forward
 * (openaimodel.py:282) when conv1's output is loop-invariant across sampling steps and cached.
 * wsets > 1: "grouped along N" like b200_conv_desc.wsets — the N images are wsets equal groups, group g uses
 * gamma / beta rows g of [wsets][C0+C1] and film / preshift advance by *_gstride floats per group (their per-image stride
 * then counts images within the group): one launch normalises the same layer of DSUnetModel's four encoders.          */
int b200_groupnorm_apply_cs(const void *x0, int C0, const double *cs0, const void *x1, int C1,
                            const double *cs1, int N, int HW, int groups, float eps, const float *gamma,
                            const float *beta, const float *film, int film_stride, int silu, void *out,
                            const float *preshift, int preshift_stride, int wsets, int film_gstride,
                            int preshift_gstride, const void *x0_lo, const void *x1_lo, int act_dtype, void *stream);
/* x0_lo / x1_lo (may be NULL): low planes of the sources (b200_conv_desc.out_lo): the value normalised is x + x_lo.   */
int b200_channel_stats(const void *x, int N, int HW, int C, double *sums, int act_dtype, void *stream);

/* ---- attention -------------------------------------------------------------------------------
 * replaces QKVAttention / QKVAttentionLegacy (openaimodel.py:537-555 / 505-521): fused
 * softmax_fp32((q*s)^T (k*s)) v with online softmax; qkv is the NHWC output of the qkv conv,
 * [B, L, 3*heads*ch].  new_order=1: q|k|v blocks of heads*ch; 0 (legacy): per head [q,k,v].
 * scale is applied to q.k (pass ch^-0.5 for the reference's ch^-0.25 on both).                 */
int b200_attention_qkv(const void *qkv, void *out, int B, int L, int heads, int ch,
                       int new_order, float scale, int act_dtype, void *stream);
/* general form: out[b, i, h*d + :] = softmax_j(scale * <q[b,i,h], k[b,j,h]>) v[b,j,h]   (fp32 softmax, exact max)
 * q [B,Lq,ldq], k [B,Lk,ldk], v [B,Lk,ldv], out [B,Lq,ldo] act; head h of q starts at channel q_off + h*q_hstride
 * (same for k, v).  Also replaces CrossAttention.forward's einsum/softmax/einsum
 * (ldm/modules/attention.py:163-194: scale = dim_head^-0.5, Lk = context length) and timm Attention inside
 * DiTBlock (UNet_DS_Diff/DiT_models.py:108).  Head dims that are multiples of 16 (32..160) run the tcgen05 flash
 * kernel; other multiples of 8 (<= 64) run the CUDA-core kernel; force_simt selects the latter for A/B tests.  */
typedef struct {
  const void *q, *k, *v;
  void *out;
  int B, Lq, Lk, heads, d;
  int ldq, ldk, ldv, ldo;
  int q_off, k_off, v_off;
  int q_hstride, k_hstride, v_hstride;
  float scale;
  int act_dtype;
  int force_simt;
  float *lse;        /* optional [B, heads, Lq] fp32: log-sum-exp of the scaled scores of every query row, saved for
                        b200_attention_bwd (training); tcgen05 kernel only (d % 16 == 0); NULL for inference            */
} b200_attn_desc;
int b200_attention(const b200_attn_desc *d, void *stream);

/* p[r][c] = softmax_c(scale * s[r][c]) : fp32 logits [rows][cols] -> act probabilities [rows][cols]
 * (torch.nn.functional.softmax(w_ * c**-0.5, dim=2), ldm/modules/diffusionmodules/model.py:189-190).              */
int b200_softmax_rows(const float *s, void *p, int64_t rows, int cols, float scale, int act_dtype, void *stream);

/* ---- KL-autoencoder edge ops (fp32 NCHW, a handful of channels) -------------------------------------------------
 * 1x1 convolution with Cin, Cout <= 16: quant_conv / post_quant_conv (ldm/models/autoencoder.py:53-54,138-147).   */
int b200_conv1x1_f32_nchw(const float *x, const float *w, const float *bias, float *out, int N, int Cin, int Cout,
                          int HW, void *stream);
/* DiagonalGaussianDistribution (ldm/modules/distributions/distributions.py:24-62) on moments [N,2C,HW]:
 * mean = first C channels, logvar = clamp(second C, -30, 20); out = mean + exp(0.5*logvar) * noise (noise NULL -> mode). */
int b200_gaussian_sample(const float *moments, const float *noise, float *out, int N, int C, int HW, void *stream);

/* ---- token-tensor ops of the transformer blocks ([rows, C] act) -------------------------------------------------
 * LayerNorm over C (gamma/beta NULL = no affine) with optional adaLN modulation y*(1+scale[b])+shift[b], b = row /
 * rows_per_batch, shift/scale fp32 with row pitch mod_stride.  Replaces nn.LayerNorm in BasicTransformerBlock
 * (ldm/modules/attention.py:319-331) and LayerNorm+modulate in DiTBlock / FinalLayer (DiT_models.py:19,107-142). */
int b200_layernorm(const void *x, const float *gamma, const float *beta, const float *shift, const float *scale,
                   int mod_stride, int rows_per_batch, int64_t rows, int C, float eps, void *out, int act_dtype,
                   void *stream);
/* out[r, c] = x[r, c] * gelu(x[r, inner + c])  (GEGLU.forward, ldm/modules/attention.py:49-56; exact erf GELU)      */
int b200_geglu(const void *x, void *out, int64_t rows, int inner, int act_dtype, void *stream);
/* DiT patch embedding input (timm PatchEmbed conv, kernel = stride = p; UNet_DS_Diff/DiT_models.py:192,243): NCHW fp32
 * -> tokens [N, (H/p)*(W/p), Kpad] act, column k = (c*p + ph)*p + pw = the column order of proj.weight.view(d, C*p*p);
 * columns >= C*p*p are zero.  The convolution itself is then b200_conv2d_igemm over the token matrix.              */
int b200_patchify(const float *x, void *out, int N, int C, int H, int W, int p, int Kpad, int act_dtype, void *stream);
/* DiT.unpatchify (DiT_models.py:220-232): tokens [N, gh*gw, ld] fp32, feature (ph*p + pw)*Cout + c -> NCHW fp32.    */
int b200_unpatchify(const float *tok, float *out, int N, int Cout, int gh, int gw, int p, int ld, void *stream);
/* out = a + b over n fp32 values (c = t_emb + y_emb, DiT_models.py:246-248).                                        */
int b200_add_f32(const float *a, const float *b, float *out, int64_t n, void *stream);
/* classifier-free guidance: out = uncond + scale * (cond - uncond) over n fp32 values, the reference's operation order
 * (DDIMSampler.p_sample_ddim, ldm/models/diffusion/ddim.py:218-219; model_wrapper, dpm_solver_pytorch.py:323-331).   */
int b200_cfg_combine(const float *uncond, const float *cond, float scale, float *out, int64_t n, void *stream);

/* ---- small dense layers on [B,K] fp32 vectors (timestep path) --------------------------------
 * replaces timestep_embedding (util.py:161-181), time_embed MLP (openaimodel.py:606-610) and
 * the per-ResBlock emb_layers Linear(SiLU(emb)) (openaimodel.py:221-228).
 * t is read from t_table[*step_ptr] when step_ptr != NULL (CUDA-graph replay), else t_table[b];
 * freqs[dim/2] = exp(-ln(max_period)*j/half), built on the host with the reference's fp32 expression. */
int b200_timestep_embedding(const float *t_table, const int *step_ptr, const float *freqs,
                            float *out, int B, int dim, void *stream);
/* out[b][o] = bias[o] + sum_k act(in[b][k]) * W[o][k];  in_silu / out_silu select the SiLU.     */
int b200_linear_f32(const float *in, const float *W, const float *bias, float *out, int B, int K,
                    int O, int in_silu, int out_silu, void *stream);
/* out = SiLU(in) over n fp32 values: the activation in front of emb_layers / adaLN_modulation Linears
 * (openaimodel.py:221-228, DiT_models.py:114-117) applied ONCE when the batch is large, instead of per output row. */
int b200_silu_f32(const float *in, float *out, int64_t n, void *stream);

/* out[0:width] = table[*step_ptr][0:width]: selects the current step's row of a table precomputed for ALL sampling
 * steps (the timestep MLP + every ResBlock's emb projection depend on t only), inside a replayed CUDA graph.      */
int b200_gather_row_f32(const float *table, const int *step_ptr, float *out, int width, void *stream);

/* ---- elementwise helpers ---------------------------------------------------------------------
 * nearest x2 upsample (Upsample.forward F.interpolate, openaimodel.py:113-117), NHWC.          */
int b200_upsample_nearest2x(const void *src, void *dst, int N, int H, int W, int C, int act_dtype,
                            void *stream);
/* Second half of a 3x3 network head with C = 1..3 output channels (self.out[-1], openaimodel.py:898-902; UNet_DS_Diff/model.py
 * `self.out`) evaluated as a 1x1 convolution to 9*C tap channels (b200_conv2d_igemm, fp32 NHWC rows of `ldp` floats,
 * tap-major: element (3r+s)*C + co) plus this gather: out[n,co,h,w] = bias[co] + sum_{r,s} taps[n,h+r-1,w+s-1,(3r+s)*C+co],
 * zero outside the image.  `out` is fp32 NCHW [N,C,H,W]; bias may be NULL.                                          */
int b200_head_gather9(const float *taps_nhwc, const float *bias, float *out_nchw, int N, int H, int W, int C, int ldp,
                      void *stream);
/* nn.AvgPool2d(2, 2) on NHWC act: Downsample(use_conv=False) and the h_upd / x_upd of ResBlock(down=True)
 * (openaimodel.py:158-160, 216-218, 265-270).  H, W even.                                                     */
int b200_avgpool2x(const void *src_nhwc, void *dst_nhwc, int N, int H, int W, int C, int act_dtype, void *stream);
/* dst = scale * (a + b + c + d)   (b,c,d may be NULL) — DS-Diff skip mean
 * (hs+hs_a+hs_al+hs_l)/4, UNet_DS_Diff/model.py:745; mean of FeatureDisentangle outputs :722.   */
int b200_add4_scale(const void *a, const void *b, const void *c, const void *d, float scale,
                    void *dst, int64_t n, int silu, int act_dtype, void *stream);
/* same, for NHWC [N,HW,C] tensors, also ACCUMULATING the per-(image, channel) statistics of dst (sums[N][C][2] fp64,
 * zero first) for the GroupNorm that consumes the DS-Diff skip mean.                                             */
int b200_add4_scale_stats(const void *a, const void *b, const void *c, const void *d, float scale, void *dst,
                          int N, int HW, int C, double *sums, const void *a_lo, const void *b_lo, const void *c_lo,
                          const void *d_lo, void *dst_lo, int act_dtype, void *stream);
/* a_lo .. d_lo (each may be NULL): low planes of the sources, added in; dst_lo (may be NULL): low plane of the result. */
/* SE_Attention (Disc_diff/guided_diffusion/unet.py:105-109): x * sigmoid(W2 relu(W1 avgpool(x))). */
int b200_se_attention(const void *x, const float *w1, const float *w2, void *out, float *scratch,
                      int N, int HW, int C, int Cr, int act_dtype, void *stream);

/* ---- fused per-step sampler update -----------------------------------------------------------
 * one kernel per step; per-step coefficients live in a device table row selected by
 * *step_ptr (so a captured CUDA graph can be replayed for every step), staged into shared
 * memory by each CTA.  All state is fp32 NCHW like the reference.
 *
 * mode B200_SAMPLER_LDM_DDIM  — DDIMSampler.p_sample_ddim  (ldm/models/diffusion/ddim.py:188-261)
 *      B200_SAMPLER_LDM_DDPM  — DDPM.p_sample/p_mean_variance/q_posterior as overridden in
 *                               trainers/trainer_ds_diff.py:717-738 (base ddpm.py:304-332)
 *      B200_SAMPLER_GD_DDIM   — GaussianDiffusion.ddim_sample
 *                               (training_project/utils/gaussian_diffusion.py:620-667, 247-351)
 *      B200_SAMPLER_GD_DDPM   — GaussianDiffusion.p_sample (:423-466)
 *      B200_SAMPLER_DPMPP_2M  — DPM-Solver++ multistep, order <= 2, data prediction: DPM_Solver.data_prediction_fn /
 *                               dpm_solver_first_update / multistep_dpm_solver_second_update
 *                               (ldm/models/diffusion/dpm_solver_new/dpm_solver_pytorch.py:435-444, 549-575, 798-835) as
 *                               driven by DPMSolverSampler.sample (dpm_solver_new/sampler.py:95-101) and by
 *                               GaussianDiffusion.dpm_solver_sample_loop (training_project/utils/gaussian_diffusion.py:
 *                               468-523).  No noise.  `hist` holds the previous step's x0 (read when the row's order is
 *                               2, always overwritten with this step's x0).  phase 0 = whole step in one launch;
 *                               phase 1 = write the raw x0 to pred_x0 only; phase 2 = dynamic thresholding
 *                               x0 = clamp(pred_x0, -s, s) / s with s = thresh[n] (dpm_solver_pytorch.py:418-427, s from
 *                               b200_abs_quantile), then the update.
 * param: 0 eps, 1 x0, 2 v.  coef row layout (B200_SAMPLER_NCOEF floats) is documented in
 * diffusion_models_dsdiff_b200/samplers.py next to the host code that fills it with the
 * reference's own float64->float32 table expressions.                                          */
#define B200_SAMPLER_LDM_DDIM 0
#define B200_SAMPLER_LDM_DDPM 1
#define B200_SAMPLER_GD_DDIM 2
#define B200_SAMPLER_GD_DDPM 3
#define B200_SAMPLER_DPMPP_2M 4
#define B200_SAMPLER_NCOEF 16
typedef struct {
  int mode, param;
  int clip_denoised;
  int learned_range;       /* GD: model_out has 2*C channels, second half = variance frac */
  int N, C, HW;            /* x is [N,C,HW]; model_out [N,C or 2C,HW]                     */
  const float *x;          /* x_t                                                         */
  const float *model_out;
  const float *noise;      /* N(0,1), same shape as x; may be NULL when every sigma is 0   */
  const float *coef;       /* [S][B200_SAMPLER_NCOEF]                                     */
  const int *step_ptr;     /* device int: row of coef for this step (NULL -> step_host)    */
  int step_host;
  float *x_prev;           /* out                                                         */
  float *pred_x0;          /* out, may be NULL                                            */
  float *hist;             /* DPMPP_2M: previous x0 [N,C,HW], in/out                       */
  const float *thresh;     /* DPMPP_2M phase 2: per-sample threshold s [N]                 */
  int phase;               /* DPMPP_2M: 0 fused step, 1 x0 only, 2 threshold + update      */
} b200_sampler_desc;
int b200_sampler_update(const b200_sampler_desc *d, void *stream);
/* out[n] = max(quantile(|x[n, :]|, q), floor_val) with torch.quantile's linear interpolation between the two
 * neighbouring order statistics (rank = q*(M-1) in fp32) — the scale of DPM-Solver++'s dynamic thresholding
 * (dpm_solver_pytorch.py:418-427).  Exact radix select, one CTA per sample.                                         */
int b200_abs_quantile(const float *x, int N, int64_t M, float q, float floor_val, float *out, void *stream);
/* out[n, :] = c1[n] * x[n, :] (+ or -) c2[n] * y[n, :] over N samples of per_sample fp32 values, coefficients gathered
 * per sample from the schedule tables by the caller: DDPM.q_sample / predict_start_from_noise /
 * predict_start_from_z_and_v / predict_eps_from_z_and_v / get_v / q_posterior mean
 * (ldm/models/diffusion/ddpm.py:284-311, 349-356; training_project/utils/gaussian_diffusion.py:184-245).           */
int b200_lincomb2(const float *x, const float *y, const float *c1, const float *c2, float *out, int N, int64_t per_sample,
                  int subtract, void *stream);
/* *step_ptr += delta (one-thread kernel; keeps the step counter on the device for graph replay). */
int b200_advance_step(int *step_ptr, int delta, void *stream);

/* ---- training step (BASELINE config 5: DiT diffusion training) ---------------------------------------
 * Everything of forward + backward + optimizer that is not a GEMM; the GEMMs run on b200_conv2d_igemm (forward and dgrad
 * with a transposed packed weight; wgrad with both operands transposed to K-major by b200_transpose16 and force_ksplit).
 * The reference gets all of this from torch autograd (loss.backward(), trainers/trainer_use_gaussian_diff.py:464-478) over
 * UNet_DS_Diff/DiT_models.py:101-142 and training_project/utils/gaussian_diffusion.py:824-983.                          */
/* dst[C, R] = src[R, C]^T, 16-bit elements; colsum (may be NULL): fp32 [C], += sum_r src[r, c] (zero it first) — the bias
 * gradient of the tensor the weight-gradient GEMM needs transposed anyway */
int b200_transpose16(const void *src, void *dst, float *colsum, int64_t R, int C, int act_dtype, void *stream);
/* out[c] += sum_r x[r, c]  (bias gradients; fp32, zero `out` first) */
int b200_colsum16(const void *x, float *out, int64_t R, int C, int act_dtype, void *stream);
/* dW[M, N] fp32 = dy^T x for TOKEN-MAJOR 16-bit dy [T, M], x [T, N] (the weight gradient of y = x W^T over all T tokens of
 * the batch) on tcgen05 with both operands read MN-major straight from those tensors — no transposed copies.  M, N multiples
 * of 64, T of 8.  b200_wgrad_splits(T, M, N) K splits (geometry only) write fp32 partial planes to `workspace`
 * (>= splits * M * N floats; may be NULL when the factor is 1) that are summed in fixed order.                           */
int b200_wgrad_splits(int64_t T, int M, int N);
int b200_linear_wgrad(const void *dy, const void *x, float *out, float *workspace, int64_t T, int M, int N, int act_dtype,
                      void *stream);
/* GELU(tanh) on a stored pre-activation, and its backward dpre = du * gelu'(pre)  (timm Mlp, DiT_models.py:111-112) */
int b200_gelu_fwd(const void *pre, void *out, int64_t n, int act_dtype, void *stream);
int b200_gelu_bwd(const void *pre, const void *du, void *dpre, int64_t n, int act_dtype, void *stream);
/* adaLN-Zero gated residual h_out = h_in + gate[n] * f (DiT_models.py:120-121; gate fp32 [N, C] rows of pitch gate_stride,
 * L token rows per sample) and its backward df = gate[n] * dh, dgate[n, c] += sum_t dh * f  (dgate zeroed by the caller) */
int b200_gated_residual(const void *h_in, const void *f, const float *gate, int gate_stride, void *h_out, int64_t rows, int L,
                        int C, int act_dtype, void *stream);
int b200_gate_bwd(const void *dh, const void *f, const float *gate, int gate_stride, void *df, float *dgate, int dgate_stride,
                  int N, int L, int C, int act_dtype, void *stream);
/* backward of a = LayerNorm(x; no affine, eps) * (1 + scale[n]) + shift[n]  (modulate(norm(x), shift, scale), :19-20, :119):
 * dx = LN-backward(da * (1 + scale)) (+ dres), dshift[n, c] += sum_t da, dscale[n, c] += sum_t da * nhat  (caller zeroes) */
int b200_ln_mod_bwd(const void *x, const void *da, const float *scale, int mod_stride, const void *dres, void *dx,
                    float *dshift, float *dscale, int dmod_stride, int N, int L, int C, float eps, int act_dtype, void *stream);
/* attention backward for self-attention over a packed qkv tensor [B, L, ld] (head h of q at q_off + h*d, ...):
 * out / dout [B, L, heads*d]; lse [B, heads, L] from b200_attention; writes dq | dk | dv into dqkv (same layout as qkv).
 * Workspaces: dsum_ws >= B*L*heads floats, dq_ws >= B*L*heads*d floats.  d in {32, 64}.                                  */
int b200_attention_bwd(const void *qkv, int ld, int q_off, int k_off, int v_off, const void *out, const void *dout,
                       const float *lse, float *dsum_ws, float *dq_ws, void *dqkv, int B, int L, int heads, int d, float scale,
                       int act_dtype, void *stream);
/* backward of DiT.unpatchify (DiT_models.py:220-232): dimg [N, Cout, gh*p, gw*p] fp32 -> dtok [N, gh*gw, ld] act, feature
 * (ph*p + pw)*Cout + c, columns >= p*p*Cout zero                                                                          */
int b200_unpatchify_bwd(const float *dimg, void *dtok, int N, int Cout, int gh, int gw, int p, int ld, int act_dtype,
                        void *stream);
/* q_sample (gaussian_diffusion.py:199-215) with per-sample coefficient rows coef[N][12] (see b200_diffusion_loss) */
int b200_q_sample(const float *x0, const float *noise, const float *coef, float *xt, int N, int per, void *stream);
/* GaussianDiffusion.training_losses, MSE / RESCALED_MSE branch (:851-977), and its gradient w.r.t. the network output:
 *   terms[N][2] = (mean_flat(Charbonnier(target, mean channels)), vb term in bits/dim x weight);
 *   d_out = d( mean_n(terms[n][0] + terms[n][1]) ) / d model_out, same shape as model_out ([N, C or 2C, HW] fp32).
 * param: 0 eps, 1 x0, 2 v.  coef[N][12] fp32 rows gathered at each sample's t on the host from the float64 tables:
 *   sqrt_ac, sqrt_1m_ac, post_mean_coef1, post_mean_coef2, post_logvar_clipped, log(beta), t == 0, vb weight
 *   (num_timesteps / 1000 for RESCALED_MSE, else 1), sqrt_recip_ac, sqrt_recipm1_ac, 0, 0.                              */
int b200_diffusion_loss(const float *model_out, const float *x_start, const float *x_t, const float *noise, const float *coef,
                        int param, int learn_sigma, float *terms, float *d_out, int N, int C, int HW, void *stream);
/* backward of the small fp32 linears of the conditioning path, y = xin W^T + b with xin = silu(x) when in_silu (N <= 64):
 *   wgrad: dW[O, I] (+)= dy^T xin, db[O] (+)= sum_n dy;   dgrad: dx[N, I] = dy W (then * silu'(x_pre) when x_pre != NULL) */
int b200_linear_f32_wgrad(const float *dy, int dy_stride, const float *x, int x_stride, int in_silu, float *dW, float *db,
                          int N, int O, int I, int accumulate, void *stream);
int b200_linear_f32_dgrad(const float *dy, int dy_stride, const float *W, const float *x_pre, float *dx, int N, int O, int I,
                          void *stream);
/* fused AdamW over a flat fp32 arena (torch.optim.AdamW semantics; step counts from 1; g is multiplied by grad_scale first:
 * 1 / world_size after a summing all-reduce).  hyper (may be NULL): device fp32 [3] = (lr, 1 - beta1^step,
 * sqrt(1 - beta2^step)) read by the kernel INSTEAD of lr / step, so a captured CUDA graph can be replayed for every step.   */
int b200_adamw(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int step, float grad_scale, const float *hyper, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DIFF_H_ */
