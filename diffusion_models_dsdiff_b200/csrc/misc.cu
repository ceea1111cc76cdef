// Small / HBM-bound kernels around the tensor-core path: API-edge layout conversion, weight repack,
// the degenerate first (Cin<=8) and last (Cout<=4) convolutions, the timestep-embedding MLP,
// nearest upsample, the DS-Diff skip mean and SE_Attention.  Reference call sites: see
// include/b200diff.h next to each entry point.
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// layout conversion (API edge)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ src, T *__restrict__ dst, int N, int C,
                                    int HW, int Cpad) {
  const long long total = (long long)N * HW * Cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long np = i / Cpad;
    const int p = (int)(np % HW);
    const int n = (int)(np / HW);
    dst[i] = c < C ? Act<T>::from_f(src[((long long)n * C + c) * HW + p]) : Act<T>::from_f(0.f);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ src, float *__restrict__ dst, int N, int C,
                                    int HW, int Csrc) {
  const long long total = (long long)N * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long nc = i / HW;
    const int c = (int)(nc % C);
    const int n = (int)(nc / C);
    dst[i] = Act<T>::to_f(src[((long long)n * HW + p) * Csrc + c]);
  }
}

template <typename T>
__global__ void pack_weight_kernel(const float *__restrict__ w, T *__restrict__ dst, int Cout, int Cin,
                                   int taps, int CinPad) {
  const long long total = (long long)taps * Cout * CinPad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % CinPad);
    const long long r = i / CinPad;
    const int co = (int)(r % Cout);
    const int tap = (int)(r / Cout);
    dst[i] = ci < Cin ? Act<T>::from_f(w[((long long)co * Cin + ci) * taps + tap]) : Act<T>::from_f(0.f);
  }
}

// ------------------------------------------------------------------------------------------
// first conv: NCHW fp32 (Cin <= 8) -> NHWC act.  Thread (row, col) of a 256-thread block owns the 8-channel column
// `col` of pixel `row` (+ k*rows): lanes of a warp write CONSECUTIVE 16-byte chunks of one pixel's channel row (fully
// coalesced 512-byte stores; the 9*Cin input taps are warp-broadcast loads), weights sit in shared memory as
// [Cin*9][Cout].  The per-(image, channel) sums for the following GroupNorm are accumulated on the fly.
// ------------------------------------------------------------------------------------------
// packed fp32x2 FMA (sm_100 FFMA2): d = a * b + d on two lanes of a 64-bit register pair
__device__ __forceinline__ void ffma2(unsigned long long &d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <typename T, bool kRegW>
__global__ void __launch_bounds__(256, 2)
conv3x3_in_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                  T *__restrict__ out, int Cin, int H, int W, int Cout, int pix_per_block, double *__restrict__ sums,
                  T *__restrict__ out_lo) {
  extern __shared__ float sw[];  // kRegW ? unused : [Cin*9][Cout]
  const int K = Cin * 9;
  if (!kRegW) {
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
      const int co = i % Cout, k = i / Cout;   // k = ci*9 + tap
      sw[i] = w[(size_t)co * K + k];
    }
    __syncthreads();
  }
  const int n = blockIdx.y;
  const int HW = H * W;
  const int ncol = Cout >> 3;
  const int rows = 256 / ncol;
  const int col = threadIdx.x % ncol;
  const int row = threadIdx.x / ncol;
  const bool active = row < rows;
  const int c0 = col * 8;
  // each thread walks a RUN of consecutive pixels, so the 3x3 input window slides (3 new loads per pixel)
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_stop = min(HW, p_begin + pix_per_block);
  const int run = (p_stop - p_begin + rows - 1) / rows;
  const int p0 = p_begin + row * run;
  const int p1 = active ? min(p_stop, p0 + run) : 0;
  unsigned long long bv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) bv[i] = pack_f32x2(bias ? bias[c0 + 2 * i] : 0.f, bias ? bias[c0 + 2 * i + 1] : 0.f);
  unsigned long long wr[kRegW ? 36 : 1];     // Cin == 1: this thread's 9 taps x 8 output channels stay in registers
  if (kRegW) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        wr[t * 4 + i] = pack_f32x2(w[(size_t)(c0 + 2 * i) * 9 + t], w[(size_t)(c0 + 2 * i + 1) * 9 + t]);
  }
  unsigned long long s2[4] = {0, 0, 0, 0}, q2[4] = {0, 0, 0, 0};
  const float *xn = x + (size_t)n * Cin * HW;
  T *on = out + (size_t)n * HW * Cout + c0;
  // one output pixel: 16-bit store (+ low plane), statistics of the stored values
  auto finish_pixel = [&](int p, const unsigned long long (&acc)[4]) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) unpack_f32x2(acc[i], f[2 * i], f[2 * i + 1]);
    const uint4 u = pack8_sat<T>(f);
    st_stream(on + (size_t)p * Cout, u);
    if (out_lo) {                       // what the 16-bit store dropped (b200_conv_desc.out_lo)
      float hf[8];
      unpack8<T>(u, hf);
#pragma unroll
      for (int i = 0; i < 8; ++i) hf[i] = f[i] - hf[i];
      st_stream(out_lo + (size_t)n * HW * Cout + c0 + (size_t)p * Cout, pack8_fast<T>(hf));
    }
    if (sums) {                         // statistics of the stored (rounded) values
      const uint32_t *uw = reinterpret_cast<const uint32_t *>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 g = unpack2<T>(uw[i]);
        const unsigned long long gg = pack_f32x2(g.x, g.y);
        ffma2(s2[i], gg, pack_f32x2(1.f, 1.f));
        ffma2(q2[i], gg, gg);
      }
    }
  };
  if constexpr (kRegW) {
    // Cin == 1 (the DS-Diff streams, once per denoising step): the run is walked image row by image row with three row
    // pointers and a sliding 3 x 3 window — three predicated loads and no 64-bit index arithmetic per pixel (the generic loop
    // below compiled to ~330 instructions per pixel for 36 useful FFMA2).  Same taps in the same order: bit-identical.
    int p = p0;
    while (p < p1) {
      const int hq = p / W, wq0 = p - hq * W;
      const int pe = min(p1, (hq + 1) * W);
      const float *r1 = xn + hq * W, *r0 = r1 - W, *r2 = r1 + W;
      const bool up = hq > 0, dn = hq + 1 < H;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;                               // column wq - 1
      if (wq0 > 0) { a0 = up ? __ldg(r0 + wq0 - 1) : 0.f; a1 = __ldg(r1 + wq0 - 1); a2 = dn ? __ldg(r2 + wq0 - 1) : 0.f; }
      float b0 = up ? __ldg(r0 + wq0) : 0.f, b1 = __ldg(r1 + wq0), b2 = dn ? __ldg(r2 + wq0) : 0.f;      // column wq
      for (int wq = wq0; p < pe; ++p, ++wq) {
        const bool rt = wq + 1 < W;
        const float c0v = (rt && up) ? __ldg(r0 + wq + 1) : 0.f, c1v = rt ? __ldg(r1 + wq + 1) : 0.f,
                    c2v = (rt && dn) ? __ldg(r2 + wq + 1) : 0.f;
        const float win[9] = {a0, b0, c0v, a1, b1, c1v, a2, b2, c2v};
        unsigned long long acc[4] = {bv[0], bv[1], bv[2], bv[3]};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const unsigned long long vv = pack_f32x2(win[t], win[t]);
#pragma unroll
          for (int i = 0; i < 4; ++i) ffma2(acc[i], vv, wr[t * 4 + i]);
        }
        finish_pixel(p, acc);
        a0 = b0; a1 = b1; a2 = b2;
        b0 = c0v; b1 = c1v; b2 = c2v;
      }
    }
  } else {
    int hq = p0 / W, wq = p0 - hq * W;
    for (int p = p0; p < p1; ++p) {
      unsigned long long acc[4] = {bv[0], bv[1], bv[2], bv[3]};
      for (int ci = 0; ci < Cin; ++ci) {
        const float *xc = xn + (size_t)ci * HW;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int hh = hq + t / 3 - 1, ww = wq + t % 3 - 1;
          const float v = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xc + hh * W + ww) : 0.f;
          const unsigned long long vv = pack_f32x2(v, v);
          const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(sw + (ci * 9 + t) * Cout + c0);
          const ulonglong2 w0 = wp[0], w1 = wp[1];
          ffma2(acc[0], vv, w0.x); ffma2(acc[1], vv, w0.y);
          ffma2(acc[2], vv, w1.x); ffma2(acc[3], vv, w1.y);
        }
      }
      finish_pixel(p, acc);
      if (++wq == W) { wq = 0; ++hq; }
    }
  }
  if (sums) {
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { unpack_f32x2(s2[i], s[2 * i], s[2 * i + 1]); unpack_f32x2(q2[i], q[2 * i], q[2 * i + 1]); }
    block_flush_col_sums(s, q, active, row, col, ncol, rows, sums + ((size_t)n * Cout + c0) * 2);
  }
}

// ------------------------------------------------------------------------------------------
// last conv: NHWC act (Cin multiple of 8) -> NCHW fp32, Cout <= 4.  One warp per output pixel
// pair... kept simple: one thread per pixel, 16-byte channel vectors through L1, weights in smem.
// ------------------------------------------------------------------------------------------
template <typename T, int COUT>
__global__ void __launch_bounds__(128)
conv3x3_out_kernel(const T *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                   float *__restrict__ out, int N, int Cin, int H, int W) {
  extern __shared__ float sw[];  // [9][Cin][COUT]
  for (int i = threadIdx.x; i < 9 * Cin * COUT; i += blockDim.x) {
    const int co = i % COUT;
    const int ci = (i / COUT) % Cin;
    const int tap = i / (COUT * Cin);
    sw[i] = w[((size_t)co * Cin + ci) * 9 + tap];
  }
  __syncthreads();
  const long long total = (long long)N * H * W;
  const long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int wq = (int)(pix % W);
  const int hq = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  float acc[COUT];
#pragma unroll
  for (int i = 0; i < COUT; ++i) acc[i] = bias ? bias[i] : 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
    const T *px = x + (((size_t)n * H + hh) * W + ww) * Cin;
    const float *wt = sw + (size_t)tap * Cin * COUT;
    for (int c = 0; c < Cin; c += 8) {
      float f[8];
      unpack8<T>(__ldg(reinterpret_cast<const uint4 *>(px + c)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = fmaf(f[i], wt[(c + i) * COUT + co], acc[co]);
    }
  }
#pragma unroll
  for (int co = 0; co < COUT; ++co) out[(((size_t)n * COUT + co) * H + hq) * W + wq] = acc[co];
}

// ------------------------------------------------------------------------------------------
// timestep path (all fp32)
// ------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float *__restrict__ t_table, const int *__restrict__ step_ptr,
                                          const float *__restrict__ freqs, float *__restrict__ out, int B,
                                          int dim) {
  const int half = dim / 2;
  const int b = blockIdx.x;
  const float t = step_ptr ? t_table[*step_ptr] : t_table[b];
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    float v = 0.f;
    if (i < 2 * half) {
      const int j = i < half ? i : i - half;
      // freqs[j] = exp(-ln(max_period) * j / half) is built on the host with the reference's own fp32
      // expression (util.py:171-173) so the angle t*freq is bit-identical; args = t * freqs (:174)
      const float arg = __fmul_rn(t, freqs[j]);
      v = i < half ? cosf(arg) : sinf(arg);
    }
    out[(size_t)b * dim + i] = v;
  }
}

__global__ void gather_row_kernel(const float *__restrict__ table, const int *__restrict__ step_ptr,
                                  float *__restrict__ out, int width) {
  const float *src = table + (size_t)(*step_ptr) * width;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < width; i += gridDim.x * blockDim.x) out[i] = src[i];
}

// out[b][o] = bias[o] + sum_k f(in[b][k]) W[o][k]; one warp per output row o, all (<= 8 per pass) b.
template <int BT>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float *__restrict__ in, const float *__restrict__ W, const float *__restrict__ bias,
                  float *__restrict__ out, int B, int K, int O, int in_silu, int out_silu) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= O) return;
  const float *wr = W + (size_t)warp * K;
  for (int b0 = 0; b0 < B; b0 += BT) {
    float acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        if (b0 + j < B) {
          float v = in[(size_t)(b0 + j) * K + k];
          if (in_silu) v = v / (1.f + expf(-v));
          acc[j] = fmaf(v, wv, acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < BT; ++j) {
      float a = acc[j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      if (lane == 0 && b0 + j < B) {
        a += bias ? bias[warp] : 0.f;
        if (out_silu) a = a / (1.f + expf(-a));
        out[(size_t)(b0 + j) * O + warp] = a;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// elementwise helpers on 16-byte vectors
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void upsample2x_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int N, int H, int W,
                                  int C8) {
  const long long total = (long long)N * (2 * H) * (2 * W) * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % (2 * W)); r /= 2 * W;
    const int ho = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    dst[i] = __ldg(src + (((long long)n * H + (ho >> 1)) * W + (wo >> 1)) * C8 + c);
  }
}

// 2x2 average pooling, NHWC act -> NHWC act (Downsample(use_conv=False) / ResBlock(down=True): nn.AvgPool2d(2, 2),
// openaimodel.py:158-160, 216-218); fp32 average, rounded once.
template <typename T>
__global__ void avgpool2x_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int N, int Ho, int Wo, int C8) {
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const long long n = r / Ho;
    const long long base = ((n * (2 * Ho) + 2 * ho) * (2 * Wo) + 2 * wo) * C8 + c;
    float acc[8] = {0}, f[8];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        unpack8<T>(__ldg(src + base + ((long long)dy * (2 * Wo) + dx) * C8), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= 0.25f;
    dst[i] = pack8_fast<T>(acc);
  }
}

template <typename T>
__global__ void add4_scale_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                  const uint4 *__restrict__ c, const uint4 *__restrict__ d, float scale,
                                  uint4 *__restrict__ dst, long long nvec, int silu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8], g[8];
    unpack8<T>(ld_stream(a + i), f);
    if (b) { unpack8<T>(ld_stream(b + i), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j]; }
    if (c) { unpack8<T>(ld_stream(c + i), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j]; }
    if (d) { unpack8<T>(ld_stream(d + i), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] *= scale;
      if (silu) f[j] = silu_f(f[j]);
    }
    dst[i] = pack8<T>(f);
  }
}

// SE_Attention: pooled[n][c] = mean_p x; y = sigmoid(W2 relu(W1 pooled)); out = x * y.
template <typename T>
__global__ void se_pool_kernel(const T *__restrict__ x, float *__restrict__ pooled, int HW, int C) {
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int p = 0; p < HW; ++p) s += Act<T>::to_f(x[((size_t)n * HW + p) * C + c]);
  pooled[(size_t)n * C + c] = s / (float)HW;
}
__global__ void se_fc_kernel(const float *__restrict__ pooled, const float *__restrict__ w1,
                             const float *__restrict__ w2, float *__restrict__ gate, int C, int Cr) {
  extern __shared__ float hid[];  // [Cr]
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(pooled[(size_t)n * C + c], w1[(size_t)j * C + c], s);
    hid[j] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(hid[j], w2[(size_t)c * Cr + j], s);
    gate[(size_t)n * C + c] = 1.f / (1.f + expf(-s));
  }
}
template <typename T>
__global__ void se_scale_kernel(const T *__restrict__ x, const float *__restrict__ gate, T *__restrict__ out,
                                int HW, int C, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((long long)HW * C));
    out[i] = Act<T>::from_f(Act<T>::to_f(x[i]) * gate[(size_t)n * C + c]);
  }
}

// row softmax: one block per row, fp32 in, 16-bit out
template <typename T>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float *__restrict__ s, T *__restrict__ p, int cols, float scale_log2e) {
  __shared__ float red[8];
  const float *row = s + (size_t)blockIdx.x * cols;
  T *out = p + (size_t)blockIdx.x * cols;
  float m = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4 *>(row + c);
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  const float mneg = -m * scale_log2e;
  float l = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4 *>(row + c);
    l += exp2f(fmaf(v.x, scale_log2e, mneg)) + exp2f(fmaf(v.y, scale_log2e, mneg)) +
         exp2f(fmaf(v.z, scale_log2e, mneg)) + exp2f(fmaf(v.w, scale_log2e, mneg));
  }
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  l = 0.f;
  for (int i = 0; i < 8; ++i) l += red[i];
  const float inv = 1.f / l;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4 *>(row + c);
    out[c + 0] = Act<T>::from_f(exp2f(fmaf(v.x, scale_log2e, mneg)) * inv);
    out[c + 1] = Act<T>::from_f(exp2f(fmaf(v.y, scale_log2e, mneg)) * inv);
    out[c + 2] = Act<T>::from_f(exp2f(fmaf(v.z, scale_log2e, mneg)) * inv);
    out[c + 3] = Act<T>::from_f(exp2f(fmaf(v.w, scale_log2e, mneg)) * inv);
  }
}

// fp32 NCHW pointwise convolution with a handful of channels (weights in registers via constant-size loops)
__global__ void conv1x1_f32_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                   const float *__restrict__ bias, float *__restrict__ out, int N, int Cin, int Cout,
                                   int HW) {
  const long long total = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), p = (int)(i % HW);
    float in[16];
    for (int c = 0; c < Cin; ++c) in[c] = x[((size_t)n * Cin + c) * HW + p];
    for (int o = 0; o < Cout; ++o) {
      float acc = bias ? bias[o] : 0.f;
      for (int c = 0; c < Cin; ++c) acc = fmaf(in[c], __ldg(w + o * Cin + c), acc);
      out[((size_t)n * Cout + o) * HW + p] = acc;
    }
  }
}

__global__ void gaussian_sample_kernel(const float *__restrict__ mom, const float *__restrict__ noise,
                                       float *__restrict__ out, int N, int C, int HW) {
  const long long total = (long long)N * C * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long chw = (long long)C * HW;
    const long long n = i / chw, r = i - n * chw;
    const float mean = mom[n * 2 * chw + r];
    float v = mean;
    if (noise) {
      const float lv = fminf(fmaxf(mom[n * 2 * chw + chw + r], -30.f), 20.f);
      v = __fadd_rn(mean, __fmul_rn(expf(__fmul_rn(0.5f, lv)), noise[i]));
    }
    out[i] = v;
  }
}

static int ew_blocks(long long n, int threads) {
  return (int)std::min<long long>(cdiv64(n, threads), (long long)num_sms() * 16);
}

}  // namespace b200

using namespace b200;
#define DISPATCH_ACT(dt, ...)                                         \
  if ((dt) == B200_F16) { using T = __half; __VA_ARGS__ }             \
  else if ((dt) == B200_BF16) { using T = __nv_bfloat16; __VA_ARGS__ } \
  else { B200_CHECK(false, "bad act_dtype %d", (dt)); }

extern "C" int b200_nchw_f32_to_nhwc(const float *src, void *dst, int N, int C, int H, int W, int Cpad,
                                     int act_dtype, void *stream) {
  B200_CHECK(src && dst && N > 0 && C > 0 && Cpad >= C, "nchw_to_nhwc: bad arguments");
  const long long total = (long long)N * H * W * Cpad;
  DISPATCH_ACT(act_dtype, nchw_to_nhwc_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              src, (T *)dst, N, C, H * W, Cpad);)
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_nhwc_to_nchw_f32(const void *src, float *dst, int N, int C, int H, int W, int Csrc,
                                     int act_dtype, void *stream) {
  B200_CHECK(src && dst && N > 0 && C > 0 && Csrc >= C, "nhwc_to_nchw: bad arguments");
  const long long total = (long long)N * H * W * C;
  DISPATCH_ACT(act_dtype, nhwc_to_nchw_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              (const T *)src, dst, N, C, H * W, Csrc);)
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_pack_conv_weight(const float *w, void *dst, int Cout, int Cin, int KH, int KW, int CinPad,
                                     int act_dtype, void *stream) {
  B200_CHECK(w && dst && Cout > 0 && Cin > 0 && CinPad >= Cin && CinPad % 8 == 0, "pack_conv_weight: bad arguments");
  const long long total = (long long)KH * KW * Cout * CinPad;
  DISPATCH_ACT(act_dtype, pack_weight_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              w, (T *)dst, Cout, Cin, KH * KW, CinPad);)
  B200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Network head with 1..3 output channels (self.out = GN, SiLU, conv3x3, openaimodel.py:898-902) as a 1x1 convolution to
// 9 * C "tap channels" on the tensor cores followed by this 9-point gather:
//   out[n, co, h, w] = bias[co] + sum_{r, s} P[n, h + r - 1, w + s - 1, (3 r + s) * C + co]      (zero outside the image)
// A 3x3 convolution to a 16-wide N tile issues 9x the MMAs of the 1x1 form for the same useful FLOPs and re-reads the
// activation 3.75x from L2; the head of cfg2 drops from 0.25 ms to the HBM time of one read of its input.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_gather9_kernel(const float *__restrict__ P, const float *__restrict__ bias, float *__restrict__ out, int N, int H, int W,
                    int C, int ldp) {
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int n = (int)(i / ((long long)W * H));
    float acc[3];
#pragma unroll
    for (int co = 0; co < 3; ++co) acc[co] = (co < C && bias) ? bias[co] : 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hh = h + r - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int ww = w + t - 1;
        if (ww < 0 || ww >= W) continue;
        const float *src = P + (((long long)n * H + hh) * W + ww) * ldp + (r * 3 + t) * C;
#pragma unroll
        for (int co = 0; co < 3; ++co)
          if (co < C) acc[co] += __ldg(src + co);
      }
    }
#pragma unroll
    for (int co = 0; co < 3; ++co)
      if (co < C) out[(((long long)n * C + co) * H + h) * W + w] = acc[co];
  }
}

extern "C" int b200_conv3x3_in(const float *x, const float *w, const float *bias, void *out, int N, int Cin,
                               int H, int W, int Cout, double *sums, void *out_lo, int act_dtype, void *stream) {
  B200_CHECK(x && w && out, "conv3x3_in: null pointer");
  B200_CHECK(Cin >= 1 && Cin <= 8, "conv3x3_in: Cin must be 1..8 (got %d)", Cin);
  B200_CHECK(Cout % 8 == 0 && Cout > 0 && Cout <= 2048, "conv3x3_in: Cout must be a multiple of 8, <= 2048");
  const size_t smem = ((size_t)Cin * 9 * Cout + Cout) * sizeof(float);
  B200_CHECK(smem <= 200 * 1024, "conv3x3_in: weights do not fit shared memory");
  const int HW = H * W;
  const int rows = std::max(1, 256 / (Cout >> 3));
  // pixel partition independent of N: the fused statistics of an image are bit-identical in any batch
  int bps = std::max(1, std::min(num_sms(), HW / (rows * 4)));
  const int ppb = cdiv(HW, bps);
  const dim3 grid(cdiv(HW, ppb), N, 1);
  DISPATCH_ACT(act_dtype, {
    if (Cin == 1) {
      // 17 KB of static shared memory per block: without a carveout hint the driver keeps the SM at its smallest shared-memory
      // configuration and ONE block per SM (ncu: occupancy limit by shared memory = 1)
      static bool carve = false;
      if (!carve) {
        B200_CUDA(cudaFuncSetAttribute(conv3x3_in_kernel<T, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
        carve = true;
      }
      conv3x3_in_kernel<T, true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, (T *)out, Cin, H, W, Cout, ppb, sums, (T *)out_lo);
    } else {
      if (smem + 16 * 1024 > 48 * 1024)     // + the static reduction buffer of block_flush_col_sums
        B200_CUDA(cudaFuncSetAttribute(conv3x3_in_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv3x3_in_kernel<T, false><<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, bias, (T *)out, Cin, H, W, Cout, ppb, sums, (T *)out_lo);
    }
  })
  B200_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int launch_conv_out(const void *x, const float *w, const float *bias, float *out, int N, int Cin, int H,
                           int W, int Cout, cudaStream_t st) {
  const size_t smem = (size_t)9 * Cin * Cout * sizeof(float);
  B200_CHECK(smem <= 200 * 1024, "conv3x3_out: weights do not fit shared memory");
  const unsigned blocks = (unsigned)cdiv64((long long)N * H * W, 128);
#define LAUNCH_OUT(CO)                                                                                     \
  {                                                                                                        \
    if (smem > 48 * 1024)                                                                                  \
      B200_CUDA(cudaFuncSetAttribute(conv3x3_out_kernel<T, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                          \
    conv3x3_out_kernel<T, CO><<<blocks, 128, smem, st>>>((const T *)x, w, bias, out, N, Cin, H, W);        \
  }
  switch (Cout) {
    case 1: LAUNCH_OUT(1) break;
    case 2: LAUNCH_OUT(2) break;
    case 3: LAUNCH_OUT(3) break;
    case 4: LAUNCH_OUT(4) break;
    default: B200_CHECK(false, "conv3x3_out: Cout must be 1..4 (got %d)", Cout);
  }
#undef LAUNCH_OUT
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_conv3x3_out(const void *x, const float *w, const float *bias, float *out, int N, int Cin,
                                int H, int W, int Cout, int act_dtype, void *stream) {
  B200_CHECK(x && w && out, "conv3x3_out: null pointer");
  B200_CHECK(Cin % 8 == 0 && Cin > 0, "conv3x3_out: Cin must be a multiple of 8");
  if (act_dtype == B200_F16) return launch_conv_out<__half>(x, w, bias, out, N, Cin, H, W, Cout, (cudaStream_t)stream);
  if (act_dtype == B200_BF16) return launch_conv_out<__nv_bfloat16>(x, w, bias, out, N, Cin, H, W, Cout, (cudaStream_t)stream);
  B200_CHECK(false, "conv3x3_out: bad act_dtype");
}

extern "C" int b200_timestep_embedding(const float *t_table, const int *step_ptr, const float *freqs, float *out,
                                       int B, int dim, void *stream) {
  B200_CHECK(t_table && freqs && out && B > 0 && dim > 0, "timestep_embedding: bad arguments");
  timestep_embedding_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(t_table, step_ptr, freqs, out, B, dim);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_softmax_rows(const float *s_, void *p, int64_t rows, int cols, float scale, int act_dtype,
                                 void *stream) {
  B200_CHECK(s_ && p && rows > 0 && cols > 0 && cols % 4 == 0, "softmax_rows: cols must be a multiple of 4");
  const float sl = scale * 1.4426950408889634f;
  DISPATCH_ACT(act_dtype, softmax_rows_kernel<T><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(s_, (T *)p, cols, sl);)
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_conv1x1_f32_nchw(const float *x, const float *w, const float *bias, float *out, int N, int Cin,
                                     int Cout, int HW, void *stream) {
  B200_CHECK(x && w && out && Cin >= 1 && Cin <= 16 && Cout >= 1 && Cout <= 16 && N > 0 && HW > 0,
             "conv1x1_f32_nchw: 1..16 channels only");
  conv1x1_f32_kernel<<<ew_blocks((long long)N * HW, 256), 256, 0, (cudaStream_t)stream>>>(x, w, bias, out, N, Cin, Cout, HW);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_gaussian_sample(const float *moments, const float *noise, float *out, int N, int C, int HW,
                                    void *stream) {
  B200_CHECK(moments && out && N > 0 && C > 0 && HW > 0, "gaussian_sample: bad arguments");
  gaussian_sample_kernel<<<ew_blocks((long long)N * C * HW, 256), 256, 0, (cudaStream_t)stream>>>(moments, noise, out, N, C, HW);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_gather_row_f32(const float *table, const int *step_ptr, float *out, int width, void *stream) {
  B200_CHECK(table && step_ptr && out && width > 0, "gather_row: bad arguments");
  gather_row_kernel<<<std::min(cdiv(width, 256), 64), 256, 0, (cudaStream_t)stream>>>(table, step_ptr, out, width);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_linear_f32(const float *in, const float *W, const float *bias, float *out, int B, int K,
                               int O, int in_silu, int out_silu, void *stream) {
  B200_CHECK(in && W && out && B > 0 && K > 0 && O > 0, "linear_f32: bad arguments");
  const int warps_per_block = 8;
  linear_f32_kernel<8><<<cdiv(O, warps_per_block), 256, 0, (cudaStream_t)stream>>>(in, W, bias, out, B, K, O,
                                                                                   in_silu, out_silu);
  B200_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) silu_f32_kernel(const float *in, float *out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = in[i];
    out[i] = v / (1.f + expf(-v));
  }
}
extern "C" int b200_silu_f32(const float *in, float *out, int64_t n, void *stream) {
  B200_CHECK(in && out && n > 0, "silu_f32: bad arguments");
  silu_f32_kernel<<<ew_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_upsample_nearest2x(const void *src, void *dst, int N, int H, int W, int C, int act_dtype,
                                       void *stream) {
  B200_CHECK(src && dst && C % 8 == 0, "upsample: C must be a multiple of 8");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  DISPATCH_ACT(act_dtype, upsample2x_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              (const uint4 *)src, (uint4 *)dst, N, H, W, C / 8);)
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_head_gather9(const float *taps, const float *bias, float *out, int N, int H, int W, int C, int ldp,
                                 void *stream) {
  B200_CHECK(taps && out && N > 0 && H > 0 && W > 0, "head_gather9: null pointer / empty tensor");
  B200_CHECK(C >= 1 && C <= 3 && ldp >= 9 * C, "head_gather9: 1..3 output channels, tap rows of at least 9*C floats");
  const long long total = (long long)N * H * W;
  head_gather9_kernel<<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(taps, bias, out, N, H, W, C, ldp);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_avgpool2x(const void *src, void *dst, int N, int H, int W, int C, int act_dtype, void *stream) {
  B200_CHECK(src && dst && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0,
             "avgpool2x: H, W must be even and C a multiple of 8");
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  DISPATCH_ACT(act_dtype, avgpool2x_kernel<T><<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              (const uint4 *)src, (uint4 *)dst, N, H / 2, W / 2, C / 8);)
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_add4_scale(const void *a, const void *b, const void *c, const void *d, float scale, void *dst,
                               int64_t n, int silu, int act_dtype, void *stream) {
  B200_CHECK(a && dst && n > 0 && n % 8 == 0, "add4_scale: element count must be a multiple of 8");
  const long long nvec = n / 8;
  DISPATCH_ACT(act_dtype, add4_scale_kernel<T><<<ew_blocks(nvec, 256), 256, 0, (cudaStream_t)stream>>>(
                              (const uint4 *)a, (const uint4 *)b, (const uint4 *)c, (const uint4 *)d, scale,
                              (uint4 *)dst, nvec, silu);)
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" int b200_se_attention(const void *x, const float *w1, const float *w2, void *out, float *scratch, int N,
                                 int HW, int C, int Cr, int act_dtype, void *stream) {
  B200_CHECK(x && w1 && w2 && out && scratch, "se_attention: null pointer");
  B200_CHECK(Cr > 0 && Cr * sizeof(float) <= 48 * 1024, "se_attention: bad reduction width");
  float *pooled = scratch;            // [N][C]
  float *gate = scratch + (size_t)N * C;  // [N][C]
  const long long total = (long long)N * HW * C;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_ACT(act_dtype, {
    se_pool_kernel<T><<<dim3(cdiv(C, 128), N), 128, 0, st>>>((const T *)x, pooled, HW, C);
    B200_LAUNCH_CHECK();
    se_fc_kernel<<<N, 256, Cr * sizeof(float), st>>>(pooled, w1, w2, gate, C, Cr);
    B200_LAUNCH_CHECK();
    se_scale_kernel<T><<<ew_blocks(total, 256), 256, 0, st>>>((const T *)x, gate, (T *)out, HW, C, total);
  })
  B200_LAUNCH_CHECK();
  return 0;
}
