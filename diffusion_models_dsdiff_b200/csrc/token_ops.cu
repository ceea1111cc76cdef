// Token-tensor kernels for the transformer blocks (SpatialTransformer, DiT): LayerNorm (+ adaLN modulate) and GEGLU over
// [rows, C] 16-bit activations.  HBM-bound, one warp per row, 16-byte vector accesses.
//
// Replaces nn.LayerNorm in BasicTransformerBlock (ldm/modules/attention.py:319-331), the affine-free LayerNorm +
// modulate(x, shift, scale) = x * (1 + scale) + shift of DiTBlock / FinalLayer (UNet_DS_Diff/DiT_models.py:19, 107-122,
// 138-142), and GEGLU.forward x * F.gelu(gate) (ldm/modules/attention.py:49-56).
#include "common.cuh"

namespace b200 {

constexpr int kLnMaxVec = 8;   // per-lane 16-byte vectors: C <= 8 * 8 * 32 = 2048

// kV = 16-byte vectors per lane (C <= 256 * kV): small rows keep few registers live, so more warps are resident and
// more bytes are in flight.  Warps are PERSISTENT: warp w handles rows w, w + W, w + 2W, ... and loads the next row's
// vectors before it reduces / normalises / stores the current one, so the three dependent phases of a row (sum, centred sum
// of squares, store) overlap with the next row's HBM latency instead of serialising with it (one short-lived row per warp
// ran at 0.30 of the HBM copy bandwidth on DiT-B's 32768 x 768 activations).
template <typename T, int kV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                 const float *__restrict__ shift, const float *__restrict__ scale, int mod_stride, int rows_per_batch,
                 long long rows, int C, float eps, T *__restrict__ out) {
  const long long stride = (long long)gridDim.x * 8;
  long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C >> 3;
  const float invC = 1.f / (float)C;
  const bool vec_mod = scale && ((((uintptr_t)scale | (uintptr_t)shift) & 15) == 0) && (mod_stride % 4 == 0);
  uint4 cur[kV], nxt[kV];
#pragma unroll
  for (int j = 0; j < kV; ++j) {
    const int vi = lane + j * 32;
    cur[j] = make_uint4(0, 0, 0, 0);
    nxt[j] = cur[j];
    if (vi < nvec) cur[j] = ld_stream(x + row * C + vi * 8);
  }
  for (; row < rows; row += stride) {
    const long long nrow = row + stride;
    if (nrow < rows) {
#pragma unroll
      for (int j = 0; j < kV; ++j) {
        const int vi = lane + j * 32;
        if (vi < nvec) nxt[j] = ld_stream(x + nrow * C + vi * 8);
      }
    }
    float v[kV][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      unpack8<T>(cur[j], v[j]);      // lanes past the row hold zeros
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[j][i];
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * invC;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      const int vi = lane + j * 32;
      if (vi < nvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; sq = fmaf(d, d, sq); }
      }
    }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * invC + eps);
    const long long b = row / rows_per_batch;
    T *orow = out + row * C;
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      const int vi = lane + j * 32;
      if (vi < nvec) {
        const int c = vi * 8;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (v[j][i] - mean) * rstd;
        if (gamma) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], gamma[c + i], beta ? beta[c + i] : 0.f);
        }
        if (scale) {
          if (vec_mod) {       // 16-byte aligned adaLN rows: two float4 loads per operand instead of eight scalar ones
            const float4 *sc = reinterpret_cast<const float4 *>(scale + b * mod_stride + c);
            const float4 *sh = reinterpret_cast<const float4 *>(shift + b * mod_stride + c);
            const float4 s0 = __ldg(sc), s1 = __ldg(sc + 1), h0 = __ldg(sh), h1 = __ldg(sh + 1);
            f[0] = fmaf(f[0], 1.f + s0.x, h0.x); f[1] = fmaf(f[1], 1.f + s0.y, h0.y);
            f[2] = fmaf(f[2], 1.f + s0.z, h0.z); f[3] = fmaf(f[3], 1.f + s0.w, h0.w);
            f[4] = fmaf(f[4], 1.f + s1.x, h1.x); f[5] = fmaf(f[5], 1.f + s1.y, h1.y);
            f[6] = fmaf(f[6], 1.f + s1.z, h1.z); f[7] = fmaf(f[7], 1.f + s1.w, h1.w);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], 1.f + scale[b * mod_stride + c + i], shift[b * mod_stride + c + i]);
          }
        }
        st_stream(orow + c, pack8_sat<T>(f));
      }
    }
#pragma unroll
    for (int j = 0; j < kV; ++j) cur[j] = nxt[j];
  }
}

__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }

template <typename T>
__global__ void __launch_bounds__(256)
geglu_kernel(const T *__restrict__ x, T *__restrict__ out, long long rows, int inner) {
  const int nvec = inner >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i - r * nvec) * 8;
    float a[8], g[8];
    unpack8<T>(ld_stream(x + r * 2 * inner + c), a);
    unpack8<T>(ld_stream(x + r * 2 * inner + inner + c), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] *= gelu_erf(g[k]);
    st_stream(out + r * inner + c, pack8<T>(a));
  }
}

// DiT patch embedding input (timm PatchEmbed = Conv2d(C, d, kernel p, stride p); UNet_DS_Diff/DiT_models.py:192,243):
// NCHW fp32 image -> tokens [N, (H/p)*(W/p), Kpad] act with k = (c*p + ph)*p + pw, the column order of
// proj.weight.view(d, C*p*p), so the convolution becomes one tensor-core GEMM over the token matrix.
template <typename T>
__global__ void __launch_bounds__(256)
patchify_kernel(const float *__restrict__ x, T *__restrict__ out, int C, int H, int W, int p, int Kpad, long long total) {
  const int gw = W / p, gh = H / p, K = C * p * p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    long long r = i / Kpad;
    const int tw = (int)(r % gw); r /= gw;
    const int th = (int)(r % gh);
    const long long n = r / gh;
    float v = 0.f;
    if (k < K) {
      const int pw = k % p, ph = (k / p) % p, c = k / (p * p);
      v = x[((n * C + c) * H + th * p + ph) * W + tw * p + pw];
    }
    out[i] = Act<T>::from_f(v);
  }
}

// DiT.unpatchify (DiT_models.py:220-232): tokens [N, gh*gw, ld] fp32 with feature index (ph*p + pw)*c_out + c
// -> NCHW fp32 image [N, c_out, gh*p, gw*p].
__global__ void __launch_bounds__(256)
unpatchify_kernel(const float *__restrict__ tok, float *__restrict__ out, int Cout, int gh, int gw, int p, int ld,
                  long long total) {
  const int H = gh * p, W = gw * p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long r = i / W;
    const int h = (int)(r % H); r /= H;
    const int c = (int)(r % Cout);
    const long long n = r / Cout;
    const int th = h / p, ph = h - th * p, tw = w / p, pw = w - tw * p;
    out[i] = tok[((n * gh + th) * gw + tw) * (long long)ld + (ph * p + pw) * Cout + c];
  }
}

// classifier-free guidance combine (ldm/models/diffusion/ddim.py:218-219): out = u + s * (c - u), reference op order, no FMA
__global__ void __launch_bounds__(256) cfg_combine_kernel(const float *u, const float *c, float s, float *out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(u[i], __fmul_rn(s, __fsub_rn(c[i], u[i])));
}

__global__ void __launch_bounds__(256) add_f32_kernel(const float *a, const float *b, float *out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

}  // namespace b200

using namespace b200;

extern "C" int b200_layernorm(const void *x, const float *gamma, const float *beta, const float *shift, const float *scale,
                              int mod_stride, int rows_per_batch, int64_t rows, int C, float eps, void *out, int act_dtype,
                              void *stream) {
  B200_CHECK(x && out && rows > 0 && C > 0 && C % 8 == 0 && C <= 8 * 32 * kLnMaxVec, "layernorm: C must be a multiple of 8, <= 2048");
  B200_CHECK((shift == nullptr) == (scale == nullptr), "layernorm: shift and scale come together");
  B200_CHECK(rows_per_batch > 0, "layernorm: rows_per_batch must be positive");
  // persistent warps: exactly one resident wave (the occupancy of this instantiation), each warp strides over the rows
#define B200_LN_LAUNCH(TT, KV)                                                                                    \
  do {                                                                                                            \
    static int per_sm = 0;                                                                                        \
    if (per_sm == 0) {                                                                                            \
      B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_kernel<TT, KV>, 256, 0));        \
      per_sm = std::max(per_sm, 1);                                                                               \
    }                                                                                                             \
    const unsigned blocks = (unsigned)std::min<long long>(cdiv64(rows, 8), (long long)num_sms() * per_sm);        \
    layernorm_kernel<TT, KV><<<blocks, 256, 0, (cudaStream_t)stream>>>((const TT *)x, gamma, beta, shift, scale,   \
                                                                       mod_stride, rows_per_batch, rows, C, eps,  \
                                                                       (TT *)out);                                \
  } while (0)
#define B200_LN_DISPATCH(TT)                     \
  do {                                           \
    if (C <= 256) B200_LN_LAUNCH(TT, 1);         \
    else if (C <= 512) B200_LN_LAUNCH(TT, 2);    \
    else if (C <= 768) B200_LN_LAUNCH(TT, 3);    \
    else if (C <= 1024) B200_LN_LAUNCH(TT, 4);   \
    else B200_LN_LAUNCH(TT, 8);                  \
  } while (0)
  if (act_dtype == B200_F16) B200_LN_DISPATCH(__half);
  else if (act_dtype == B200_BF16) B200_LN_DISPATCH(__nv_bfloat16);
  else B200_CHECK(false, "layernorm: bad act_dtype");
#undef B200_LN_DISPATCH
#undef B200_LN_LAUNCH
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_geglu(const void *x, void *out, int64_t rows, int inner, int act_dtype, void *stream) {
  B200_CHECK(x && out && rows > 0 && inner > 0 && inner % 8 == 0, "geglu: inner width must be a multiple of 8");
  const long long total = rows * (inner / 8);
  const int blocks = (int)std::min<long long>(cdiv64(total, 256), (long long)num_sms() * 16);
  if (act_dtype == B200_F16)
    geglu_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __half *)x, (__half *)out, rows, inner);
  else if (act_dtype == B200_BF16)
    geglu_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)out, rows, inner);
  else
    B200_CHECK(false, "geglu: bad act_dtype");
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_patchify(const float *x, void *out, int N, int C, int H, int W, int p, int Kpad, int act_dtype,
                             void *stream) {
  B200_CHECK(x && out && N > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0, "patchify: H, W must be multiples of p");
  B200_CHECK(Kpad >= C * p * p && Kpad % 8 == 0, "patchify: Kpad must be a multiple of 8 >= C*p*p");
  const long long total = (long long)N * (H / p) * (W / p) * Kpad;
  const int blocks = (int)std::min<long long>(cdiv64(total, 256), (long long)num_sms() * 16);
  if (act_dtype == B200_F16)
    patchify_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__half *)out, C, H, W, p, Kpad, total);
  else if (act_dtype == B200_BF16)
    patchify_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16 *)out, C, H, W, p, Kpad, total);
  else
    B200_CHECK(false, "patchify: bad act_dtype");
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_unpatchify(const float *tok, float *out, int N, int Cout, int gh, int gw, int p, int ld, void *stream) {
  B200_CHECK(tok && out && N > 0 && Cout > 0 && gh > 0 && gw > 0 && p > 0 && ld >= p * p * Cout, "unpatchify: bad geometry");
  const long long total = (long long)N * Cout * gh * p * gw * p;
  const int blocks = (int)std::min<long long>(cdiv64(total, 256), (long long)num_sms() * 16);
  unpatchify_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tok, out, Cout, gh, gw, p, ld, total);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_add_f32(const float *a, const float *b, float *out, int64_t n, void *stream) {
  B200_CHECK(a && b && out && n > 0, "add_f32: bad arguments");
  const int blocks = (int)std::min<long long>(cdiv64(n, 256), (long long)num_sms() * 16);
  add_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_cfg_combine(const float *uncond, const float *cond, float scale, float *out, int64_t n, void *stream) {
  B200_CHECK(uncond && cond && out && n > 0, "cfg_combine: bad arguments");
  const int blocks = (int)std::min<long long>(cdiv64(n, 256), (long long)num_sms() * 16);
  cfg_combine_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(uncond, cond, scale, out, n);
  B200_LAUNCH_CHECK();
  return 0;
}
