// Fused self-attention with online softmax (never materialises the L x L matrix).
// Replaces QKVAttention / QKVAttentionLegacy, ldm/modules/diffusionmodules/openaimodel.py:537-555 /
// 505-521:  w = softmax_fp32((q*s)^T (k*s)),  a = w v,  s = ch^-1/4.
//
// v1: fp32 CUDA-core kernel — one query per thread, K/V tiles staged in shared memory as fp32 and
// read with broadcast LDS.128.  Attention is 0.2-2 % of the network FLOPs; the
// tcgen05 version is the planned replacement once the convolutions are at roofline.
#include "common.cuh"

namespace b200 {

constexpr int kAttnThreads = 128;
constexpr int kKT = 32;  // keys per tile

template <typename T, int CH>
__global__ void __launch_bounds__(kAttnThreads)
attn_kernel(const T *__restrict__ qkv, T *__restrict__ out, int L, int heads, int new_order, float scale_log2e) {
  __shared__ __align__(16) float sK[kKT][CH];
  __shared__ __align__(16) float sV[kKT][CH];
  const int h = blockIdx.y, b = blockIdx.z;
  const int C = heads * CH;
  const int C3 = 3 * C;
  const int qoff = new_order ? h * CH : h * 3 * CH;
  const int koff = new_order ? C + h * CH : h * 3 * CH + CH;
  const int voff = new_order ? 2 * C + h * CH : h * 3 * CH + 2 * CH;
  const int qi = blockIdx.x * kAttnThreads + threadIdx.x;
  const bool valid = qi < L;
  const T *base = qkv + (size_t)b * L * C3;

  float q[CH], acc[CH];
#pragma unroll
  for (int c = 0; c < CH; c += 8) {
    float f[8];
    if (valid) unpack8<T>(*reinterpret_cast<const uint4 *>(base + (size_t)qi * C3 + qoff + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      q[c + i] = valid ? f[i] * scale_log2e : 0.f;
      acc[c + i] = 0.f;
    }
  }
  float m = -INFINITY, l = 0.f;

  for (int kt = 0; kt < L; kt += kKT) {
    __syncthreads();
    constexpr int kVecPerRow = CH / 8;
    for (int i = threadIdx.x; i < kKT * kVecPerRow * 2; i += kAttnThreads) {
      const int which = i / (kKT * kVecPerRow);
      const int r = (i % (kKT * kVecPerRow)) / kVecPerRow;
      const int cv = (i % kVecPerRow) * 8;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (kt + r < L)
        unpack8<T>(*reinterpret_cast<const uint4 *>(base + (size_t)(kt + r) * C3 + (which ? voff : koff) + cv), f);
      float *dst = which ? &sV[r][cv] : &sK[r][cv];
      *reinterpret_cast<float4 *>(dst) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4 *>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    const int nk = min(kKT, L - kt);
    float s[kKT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < kKT; ++j) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < CH; c += 4) {
        const float4 kv = *reinterpret_cast<const float4 *>(&sK[j][c]);
        d = fmaf(q[c], kv.x, d); d = fmaf(q[c + 1], kv.y, d);
        d = fmaf(q[c + 2], kv.z, d); d = fmaf(q[c + 3], kv.w, d);
      }
      s[j] = j < nk ? d : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = exp2f(m - m_new);   // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] *= corr;
#pragma unroll
    for (int j = 0; j < kKT; ++j) {
      const float p = exp2f(s[j] - m_new);
      l += p;
#pragma unroll
      for (int c = 0; c < CH; c += 4) {
        const float4 vv = *reinterpret_cast<const float4 *>(&sV[j][c]);
        acc[c] = fmaf(p, vv.x, acc[c]); acc[c + 1] = fmaf(p, vv.y, acc[c + 1]);
        acc[c + 2] = fmaf(p, vv.z, acc[c + 2]); acc[c + 3] = fmaf(p, vv.w, acc[c + 3]);
      }
    }
    m = m_new;
  }
  if (valid) {
    const float inv = 1.f / l;
    T *o = out + ((size_t)b * L + qi) * C + h * CH;
#pragma unroll
    for (int c = 0; c < CH; c += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = acc[c + i] * inv;
      *reinterpret_cast<uint4 *>(o + c) = pack8<T>(f);
    }
  }
}

template <typename T>
static int launch_attn(const void *qkv, void *out, int B, int L, int heads, int ch, int new_order, float scale,
                       cudaStream_t st) {
  const dim3 grid(cdiv(L, kAttnThreads), heads, B);
  const float sl = scale * 1.4426950408889634f;
#define ATTN_CASE(CHV)                                                                                       \
  case CHV:                                                                                                  \
    attn_kernel<T, CHV><<<grid, kAttnThreads, 0, st>>>((const T *)qkv, (T *)out, L, heads, new_order, sl); \
    break;
  switch (ch) {
    ATTN_CASE(8) ATTN_CASE(16) ATTN_CASE(24) ATTN_CASE(32) ATTN_CASE(40) ATTN_CASE(48) ATTN_CASE(64)
    default: B200_CHECK(false, "attention: unsupported head dim %d", ch);
  }
#undef ATTN_CASE
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200

extern "C" int b200_attention_qkv(const void *qkv, void *out, int B, int L, int heads, int ch, int new_order,
                                  float scale, int act_dtype, void *stream) {
  using namespace b200;
  B200_CHECK(qkv && out && B > 0 && L > 0 && heads > 0, "attention: bad arguments");
  if (act_dtype == B200_F16) return launch_attn<__half>(qkv, out, B, L, heads, ch, new_order, scale, (cudaStream_t)stream);
  if (act_dtype == B200_BF16) return launch_attn<__nv_bfloat16>(qkv, out, B, L, heads, ch, new_order, scale, (cudaStream_t)stream);
  B200_CHECK(false, "attention: bad act_dtype");
}
