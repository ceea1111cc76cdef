// GroupNorm statistics + fused GroupNorm(+FiLM)(+SiLU) apply over NHWC 16-bit activations.
// HBM-bound: every access is a coalesced 16-byte vector; reductions are warp shuffles + shared
// fp64 atomics (one global fp64 atomic pair per (block, group)).
//
// Replaces GroupNorm32 -> SiLU (ldm/modules/diffusionmodules/util.py:209-226; openaimodel.py:
// 205-209, 229-236, 898-902), the FiLM form (openaimodel.py:276-280) and Normalize
// (ldm/modules/attention.py:88-89).  The input may be the channel concatenation [x0 | x1]
// (th.cat([h, hs.pop()], 1), openaimodel.py:952) which is never materialised un-normalised.
#include "common.cuh"

namespace b200 {

constexpr int kNormThreads = 256;
constexpr int kMaxGroups = 64;

// Each thread owns one 8-channel column of one source and walks pixels with a fixed stride, so
// per-channel partial sums stay in registers.
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
gn_stats_kernel(const T *__restrict__ x0, int C0, const T *__restrict__ x1, int C1, int HW,
                int groups, int pix_per_block, double *__restrict__ stats) {
  __shared__ double s_sum[kMaxGroups], s_sq[kMaxGroups];
  const int n = blockIdx.y;
  const int C = C0 + C1;
  const int cpg = C / groups;
  for (int i = threadIdx.x; i < groups; i += blockDim.x) { s_sum[i] = 0.0; s_sq[i] = 0.0; }
  __syncthreads();
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  for (int src = 0; src < 2; ++src) {
    const T *x = src == 0 ? x0 : x1;
    const int Cs = src == 0 ? C0 : C1;
    const int cbase = src == 0 ? 0 : C0;
    if (Cs == 0) continue;
    const int ncol = Cs >> 3;
    const int rows = kNormThreads / ncol;   // pixels per sweep
    const int col = threadIdx.x % ncol;
    const int row = threadIdx.x / ncol;
    if (row < rows) {
      float s[8] = {0}, q[8] = {0};
      const T *base = x + (size_t)n * HW * Cs + col * 8;
      int p = p_begin + row;
      // 4 independent 16-byte loads in flight per thread
      for (; p + 3 * rows < p_end; p += 4 * rows) {
        uint4 u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ld_stream(base + (size_t)(p + j * rows) * Cs);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float f[8];
          unpack8<T>(u[j], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
        }
      }
      for (; p < p_end; p += rows) {
        float f[8];
        unpack8<T>(ld_stream(base + (size_t)p * Cs), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
      }
      // fold the 8 channels into (at most a few) groups before touching shared memory
      int g_cur = (cbase + col * 8) / cpg;
      float gs = 0.f, gq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int g = (cbase + col * 8 + i) / cpg;
        if (g != g_cur) {
          atomicAdd(&s_sum[g_cur], (double)gs);
          atomicAdd(&s_sq[g_cur], (double)gq);
          g_cur = g; gs = 0.f; gq = 0.f;
        }
        gs += s[i]; gq += q[i];
      }
      atomicAdd(&s_sum[g_cur], (double)gs);
      atomicAdd(&s_sq[g_cur], (double)gq);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups; i += blockDim.x) {
    atomicAdd(&stats[((size_t)n * groups + i) * 2 + 0], s_sum[i]);
    atomicAdd(&stats[((size_t)n * groups + i) * 2 + 1], s_sq[i]);
  }
}

// Per-(image, channel) sum and sum of squares -> sums[N][C][2] (fp64, accumulated: zero it first).  Same thread
// mapping as gn_stats_kernel; used for tensors whose producer has no fused statistics epilogue.
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
channel_stats_kernel(const T *__restrict__ x, int C, int HW, int pix_per_block, double *__restrict__ sums) {
  const int n = blockIdx.y;
  const int ncol = C >> 3;
  const int rows = kNormThreads / ncol;
  const int col = threadIdx.x % ncol;
  const int row = threadIdx.x / ncol;
  const bool active = row < rows;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = active ? min(HW, p_begin + pix_per_block) : 0;
  float s[8] = {0}, q[8] = {0};
  const T *base = x + (size_t)n * HW * C + col * 8;
  int p = p_begin + row;
  for (; p + 3 * rows < p_end; p += 4 * rows) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = ld_stream(base + (size_t)(p + j * rows) * C);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
      unpack8<T>(u[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
    }
  }
  for (; p < p_end; p += rows) {
    float f[8];
    unpack8<T>(ld_stream(base + (size_t)p * C), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
  }
  block_flush_col_sums(s, q, active, row, col, ncol, rows, sums + ((size_t)n * C + col * 8) * 2);
}

template <typename T, bool kLo>
__device__ __forceinline__ void
add4_stats_body(const T *__restrict__ a, const T *__restrict__ b, const T *__restrict__ c, const T *__restrict__ d,
                float scale, T *__restrict__ dst, int C, int HW, int pix_per_block, double *__restrict__ sums,
                const T *__restrict__ a_lo, const T *__restrict__ b_lo, const T *__restrict__ c_lo,
                const T *__restrict__ d_lo, T *__restrict__ dst_lo) {
  const int n = blockIdx.y;
  const int ncol = C >> 3;
  const int rows = kNormThreads / ncol;
  const int col = threadIdx.x % ncol;
  const int row = threadIdx.x / ncol;
  const bool active = row < rows;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = active ? min(HW, p_begin + pix_per_block) : 0;
  float s[8] = {0}, q[8] = {0};
  const size_t base = (size_t)n * HW * C + col * 8;
  const int nin = 1 + (b != nullptr) + (c != nullptr) + (d != nullptr);
  const T *src[4] = {a, b, c, d};
  const T *los[4] = {a_lo, b_lo, c_lo, d_lo};
  (void)los;
  auto finish = [&](float (&f)[8], size_t off) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] *= scale;
    const uint4 u = pack8_fast<T>(f);     // a mean of stored 16-bit values cannot overflow
    st_stream(dst + off, u);
    if constexpr (kLo) {
      if (dst_lo) {
        float hf[8];
        unpack8<T>(u, hf);
#pragma unroll
        for (int i = 0; i < 8; ++i) hf[i] = f[i] - hf[i];
        st_stream(dst_lo + off, pack8_fast<T>(hf));
      }
    }
    unpack8<T>(u, f);                     // statistics of the stored (rounded) values, like channel_stats on dst
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
  };
  int p = p_begin + row;
  for (; p + rows < p_end; p += 2 * rows) {      // two pixels per trip: up to 8 independent 16-byte loads in flight
    const size_t off0 = base + (size_t)p * C, off1 = off0 + (size_t)rows * C;
    uint4 u0[4], u1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nin) { u0[j] = ld_stream(src[j] + off0); u1[j] = ld_stream(src[j] + off1); }
    float f0[8] = {0}, f1[8] = {0}, g[8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nin) {
        unpack8<T>(u0[j], g);
#pragma unroll
        for (int i = 0; i < 8; ++i) f0[i] += g[i];
        unpack8<T>(u1[j], g);
#pragma unroll
        for (int i = 0; i < 8; ++i) f1[i] += g[i];
      }
    if constexpr (kLo) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nin && los[j]) {
          const uint4 l0 = ld_stream(los[j] + off0), l1 = ld_stream(los[j] + off1);
          unpack8<T>(l0, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) f0[i] += g[i];
          unpack8<T>(l1, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) f1[i] += g[i];
        }
    }
    finish(f0, off0);
    finish(f1, off1);
  }
  for (; p < p_end; p += rows) {
    const size_t off = base + (size_t)p * C;
    float f[8] = {0}, g[8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nin) {
        unpack8<T>(ld_stream(src[j] + off), g);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += g[i];
        if constexpr (kLo) {
          if (los[j]) {
            unpack8<T>(ld_stream(los[j] + off), g);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += g[i];
          }
        }
      }
    finish(f, off);
  }
  block_flush_col_sums(s, q, active, row, col, ncol, rows, sums + ((size_t)n * C + col * 8) * 2);
}

// dst = scale * (a + b + c + d) with the per-(image, channel) statistics of dst accumulated on the fly (DS-Diff skip mean,
// UNet_DS_Diff/model.py:745, whose result feeds the decoder's GroupNorm).  Thread mapping as channel_stats_kernel.
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
add4_stats_kernel(const T *__restrict__ a, const T *__restrict__ b, const T *__restrict__ c, const T *__restrict__ d,
                  float scale, T *__restrict__ dst, int C, int HW, int pix_per_block, double *__restrict__ sums) {
  add4_stats_body<T, false>(a, b, c, d, scale, dst, C, HW, pix_per_block, sums, nullptr, nullptr, nullptr, nullptr, nullptr);
}

// the same with LOW planes (b200_conv_desc.out_lo) of some sources added in and / or a low plane of the result written
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
add4_stats_lo_kernel(const T *__restrict__ a, const T *__restrict__ b, const T *__restrict__ c, const T *__restrict__ d,
                     float scale, T *__restrict__ dst, int C, int HW, int pix_per_block, double *__restrict__ sums,
                     const T *__restrict__ a_lo, const T *__restrict__ b_lo, const T *__restrict__ c_lo,
                     const T *__restrict__ d_lo, T *__restrict__ dst_lo) {
  add4_stats_body<T, true>(a, b, c, d, scale, dst, C, HW, pix_per_block, sums, a_lo, b_lo, c_lo, d_lo, dst_lo);
}

// y = silu?( (x - mean_g) * rstd_g * gamma_c + beta_c ) [ * (1 + scale_nc) + shift_nc ]
template <typename T, bool kLo>
__global__ void __launch_bounds__(kNormThreads, 4)
gn_apply_kernel(const T *__restrict__ x0, int C0, const T *__restrict__ x1, int C1, int HW,
                int groups, float eps, int pix_per_block, const double *__restrict__ stats,
                const double *__restrict__ cs0, const double *__restrict__ cs1,
                const float *__restrict__ gamma, const float *__restrict__ beta,
                const float *__restrict__ film, int film_stride, int silu, T *__restrict__ out,
                const float *__restrict__ preshift, int preshift_stride, int pdl, int n_per_set, int film_gstride,
                int preshift_gstride, const T *__restrict__ x0_lo, const T *__restrict__ x1_lo) {
  // kLo: a source may come with a LOW plane (b200_conv_desc.out_lo): the value normalised is x + x_lo
  if (pdl) {      // programmatic dependent launch (see igemm_conv.cu): wait for the producer, release the consumer
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  // preshift (optional, first source only): the tensor normalised is x0 + preshift[n, c] — a per-(image, channel)
  // constant added on the fly (the timestep-embedding add of a ResBlock whose conv1 output is cached across steps); the
  // producer's sums of x0 are corrected analytically: sum' = sum + HW s, sq' = sq + 2 s sum + HW s^2
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  const int n = blockIdx.y;
  const int C = C0 + C1;
  const int cpg = C / groups;
  const double cnt = (double)HW * (double)cpg;
  {   // grouped along N (see b200_groupnorm_apply_cs): image n belongs to weight set n / n_per_set
    const int wset = n / n_per_set, nl = n - wset * n_per_set;
    gamma += (size_t)wset * C;
    beta += (size_t)wset * C;
    if (film) film += (size_t)wset * film_gstride + (size_t)nl * film_stride;
    if (preshift) preshift += (size_t)wset * preshift_gstride + (size_t)nl * preshift_stride;
  }
  // group statistics: tpg threads per group sum the producers' per-channel fp64 sums, then a shuffle tree (the serial
  // 30..60-channel loop of one thread per group was the launch floor of the small-map GroupNorms)
  int tpg = 1;
  while (tpg * 2 * groups <= kNormThreads && tpg < 32) tpg *= 2;
  {
    const int g = threadIdx.x / tpg, sub = threadIdx.x % tpg;
    double sum = 0.0, sq = 0.0;
    if (g < groups) {
      if (stats) {
        if (sub == 0) {
          sum = stats[((size_t)n * groups + g) * 2 + 0];
          sq = stats[((size_t)n * groups + g) * 2 + 1];
        }
      } else {   // per-channel sums written by the producers' epilogues (conv / add4 / channel_stats)
        for (int c = g * cpg + sub; c < (g + 1) * cpg; c += tpg) {
          const double *src = c < C0 ? cs0 + ((size_t)n * C0 + c) * 2 : cs1 + ((size_t)n * C1 + (c - C0)) * 2;
          double cs = src[0], cq = src[1];
          if (preshift && c < C0) {
            const double sh = (double)preshift[c];
            cq += 2.0 * sh * cs + (double)HW * sh * sh;
            cs += (double)HW * sh;
          }
          sum += cs; sq += cq;
        }
      }
    }
    for (int o = tpg >> 1; o > 0; o >>= 1) {      // tpg divides 32: the tree stays inside one warp
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (g < groups && sub == 0) {
      const double m = sum / cnt;
      double var = sq / cnt - m * m;
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)m;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  const int ncol = C >> 3;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  const int cpp = min(ncol, kNormThreads);   // 8-channel columns per pass (more than 2048 channels: several passes)
  const int rows = kNormThreads / cpp;
  const int row = threadIdx.x / cpp;
  for (int col = threadIdx.x % cpp; col < ncol && row < rows; col += cpp) {
    const int c = col * 8;
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c + i) / cpg;
      const float ga = gamma[c + i] * s_rstd[g];
      float aa = ga;
      float bb = beta[c + i] - s_mean[g] * ga;
      if (preshift && c + i < C0) bb = fmaf(ga, preshift[c + i], bb);
      if (film) {
        const float sc = 1.f + film[c + i];
        const float sh = film[C + c + i];
        aa *= sc;
        bb = bb * sc + sh;
      }
      a[i] = aa; b[i] = bb;
    }
    const bool from0 = c < C0;
    const T *src = from0 ? x0 + (size_t)n * HW * C0 + c : x1 + (size_t)n * HW * C1 + (c - C0);
    const int Cs = from0 ? C0 : C1;
    T *dst = out + (size_t)n * HW * C + c;
    int p = p_begin + row;
    if constexpr (kLo) {
      const T *lo = from0 ? x0_lo : x1_lo;
      if (lo != nullptr) {
        lo += (size_t)n * HW * Cs + (from0 ? c : c - C0);
        for (; p + rows < p_end; p += 2 * rows) {
          uint4 u[2], ul[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            u[j] = ld_stream(src + (size_t)(p + j * rows) * Cs);
            ul[j] = ld_stream(lo + (size_t)(p + j * rows) * Cs);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float f[8], fl[8];
            unpack8<T>(u[j], f);
            unpack8<T>(ul[j], fl);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i] + fl[i], a[i], b[i]);
            if (silu) silu8(f);
            st_stream(dst + (size_t)(p + j * rows) * C, pack8_fast<T>(f));
          }
        }
        for (; p < p_end; p += rows) {
          float f[8], fl[8];
          unpack8<T>(ld_stream(src + (size_t)p * Cs), f);
          unpack8<T>(ld_stream(lo + (size_t)p * Cs), fl);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i] + fl[i], a[i], b[i]);
          if (silu) silu8(f);
          st_stream(dst + (size_t)p * C, pack8_fast<T>(f));
        }
        continue;
      }
    }
    for (; p + 3 * rows < p_end; p += 4 * rows) {
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = ld_stream(src + (size_t)(p + j * rows) * Cs);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8<T>(u[j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], a[i], b[i]);
        if (silu) silu8(f);
        st_stream(dst + (size_t)(p + j * rows) * C, pack8_fast<T>(f));
      }
    }
    for (; p < p_end; p += rows) {
      float f[8];
      unpack8<T>(ld_stream(src + (size_t)p * Cs), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], a[i], b[i]);
      if (silu) silu8(f);
      st_stream(dst + (size_t)p * C, pack8_fast<T>(f));
    }
  }
}

// Kernels that PRODUCE statistics (fp32 per-thread partial sums) split an image's pixels the same way whatever the batch
// size, so an image's sums — and everything downstream — are bit-identical in any batch.
static void stats_grid(int N, int HW, int ncol, dim3 *grid, int *pix_per_block) {
  const int rows = std::max(1, kNormThreads / std::max(ncol, 1));
  const int bps = std::max(1, std::min(num_sms(), HW / (rows * 4)));
  *pix_per_block = cdiv(HW, bps);
  *grid = dim3(cdiv(HW, *pix_per_block), N, 1);
}

static void norm_grid(int N, int HW, int ncol, dim3 *grid, int *pix_per_block) {
  const int rows = std::max(1, kNormThreads / std::max(ncol, 1));
  int target_blocks = std::max(1, (num_sms() * 4) / std::max(N, 1));   // one resident wave (4 CTAs of 256 threads per SM)
  // at least 4 sweeps of `rows` pixels per block so the unrolled loop body is used
  int max_blocks = std::max(1, HW / (rows * 4));
  int bps = std::min(target_blocks, max_blocks);
  *pix_per_block = cdiv(HW, bps);
  *grid = dim3(cdiv(HW, *pix_per_block), N, 1);
}

template <typename T>
static int gn_stats(const void *x0, int C0, const void *x1, int C1, int N, int HW, int groups,
                    double *stats, cudaStream_t st) {
  dim3 grid;
  int ppb;
  stats_grid(N, HW, std::max(C0, C1) >> 3, &grid, &ppb);
  gn_stats_kernel<T><<<grid, kNormThreads, 0, st>>>((const T *)x0, C0, (const T *)x1, C1, HW, groups,
                                                   ppb, stats);
  B200_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int gn_apply(const void *x0, int C0, const void *x1, int C1, int N, int HW, int groups,
                    float eps, const double *stats, const double *cs0, const double *cs1, const float *gamma,
                    const float *beta, const float *film, int film_stride, int silu, void *out, cudaStream_t st,
                    const float *preshift = nullptr, int preshift_stride = 0, int wsets = 1, int film_gstride = 0,
                    int preshift_gstride = 0, const void *x0_lo = nullptr, const void *x1_lo = nullptr) {
  wsets = std::max(1, wsets);
  B200_CHECK(N % wsets == 0, "groupnorm: N (%d) must be a multiple of the number of weight sets (%d)", N, wsets);
  dim3 grid;
  int ppb;
  norm_grid(N, HW, std::min((C0 + C1) >> 3, kNormThreads), &grid, &ppb);
  static const bool pdl = getenv("B200DIFF_PDL") != nullptr && getenv("B200DIFF_PDL")[0] == '1';
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kNormThreads, 1, 1);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (x0_lo || x1_lo)
    B200_CUDA(cudaLaunchKernelEx(&cfg, gn_apply_kernel<T, true>, (const T *)x0, C0, (const T *)x1, C1, HW, groups, eps, ppb, stats,
                                 cs0, cs1, gamma, beta, film, film_stride, silu, (T *)out, preshift, preshift_stride, (int)pdl,
                                 N / wsets, film_gstride, preshift_gstride, (const T *)x0_lo, (const T *)x1_lo));
  else
    B200_CUDA(cudaLaunchKernelEx(&cfg, gn_apply_kernel<T, false>, (const T *)x0, C0, (const T *)x1, C1, HW, groups, eps, ppb, stats,
                                 cs0, cs1, gamma, beta, film, film_stride, silu, (T *)out, preshift, preshift_stride, (int)pdl,
                                 N / wsets, film_gstride, preshift_gstride, (const T *)nullptr, (const T *)nullptr));
  B200_LAUNCH_CHECK();
  return 0;
}

static int check_gn(int C0, int C1, int N, int HW, int groups) {
  B200_CHECK(C0 > 0 && C0 % 8 == 0 && C1 % 8 == 0 && C1 >= 0, "groupnorm: channels must be multiples of 8");
  B200_CHECK(groups > 0 && groups <= kMaxGroups && (C0 + C1) % groups == 0, "groupnorm: bad group count");
  B200_CHECK(N > 0 && HW > 0, "groupnorm: empty tensor");
  B200_CHECK(C0 + C1 <= 64 * kNormThreads, "groupnorm: at most %d channels", 64 * kNormThreads);
  return 0;
}

}  // namespace b200

extern "C" int b200_groupnorm_stats(const void *x0, int C0, const void *x1, int C1, int N, int HW,
                                    int groups, double *stats, int act_dtype, void *stream) {
  using namespace b200;
  if (x1 == nullptr) C1 = 0;
  if (int rc = check_gn(C0, C1, N, HW, groups)) return rc;
  if (act_dtype == B200_F16) return gn_stats<__half>(x0, C0, x1, C1, N, HW, groups, stats, (cudaStream_t)stream);
  if (act_dtype == B200_BF16) return gn_stats<__nv_bfloat16>(x0, C0, x1, C1, N, HW, groups, stats, (cudaStream_t)stream);
  B200_CHECK(false, "groupnorm: bad act_dtype");
}

extern "C" int b200_groupnorm_apply(const void *x0, int C0, const void *x1, int C1, int N, int HW,
                                    int groups, float eps, const double *stats, const float *gamma,
                                    const float *beta, const float *film, int film_stride, int silu,
                                    void *out, int act_dtype, void *stream) {
  using namespace b200;
  if (x1 == nullptr) C1 = 0;
  if (int rc = check_gn(C0, C1, N, HW, groups)) return rc;
  if (act_dtype == B200_F16)
    return gn_apply<__half>(x0, C0, x1, C1, N, HW, groups, eps, stats, nullptr, nullptr, gamma, beta, film, film_stride, silu, out, (cudaStream_t)stream);
  if (act_dtype == B200_BF16)
    return gn_apply<__nv_bfloat16>(x0, C0, x1, C1, N, HW, groups, eps, stats, nullptr, nullptr, gamma, beta, film, film_stride, silu, out, (cudaStream_t)stream);
  B200_CHECK(false, "groupnorm: bad act_dtype");
}

extern "C" int b200_groupnorm_apply_cs(const void *x0, int C0, const double *cs0, const void *x1, int C1,
                                       const double *cs1, int N, int HW, int groups, float eps, const float *gamma,
                                       const float *beta, const float *film, int film_stride, int silu, void *out,
                                       const float *preshift, int preshift_stride, int wsets, int film_gstride,
                                       int preshift_gstride, const void *x0_lo, const void *x1_lo, int act_dtype,
                                       void *stream) {
  using namespace b200;
  if (x1 == nullptr) C1 = 0;
  if (int rc = check_gn(C0, C1, N, HW, groups)) return rc;
  B200_CHECK(cs0 != nullptr && (C1 == 0 || cs1 != nullptr), "groupnorm_apply_cs: missing channel sums");
  if (C1 == 0) x1_lo = nullptr;
  B200_CHECK(!(preshift && x0_lo), "groupnorm_apply_cs: preshift and a low plane of x0 cannot be combined");
  if (act_dtype == B200_F16)
    return gn_apply<__half>(x0, C0, x1, C1, N, HW, groups, eps, nullptr, cs0, cs1, gamma, beta, film, film_stride, silu, out, (cudaStream_t)stream, preshift, preshift_stride, wsets, film_gstride, preshift_gstride, x0_lo, x1_lo);
  if (act_dtype == B200_BF16)
    return gn_apply<__nv_bfloat16>(x0, C0, x1, C1, N, HW, groups, eps, nullptr, cs0, cs1, gamma, beta, film, film_stride, silu, out, (cudaStream_t)stream, preshift, preshift_stride, wsets, film_gstride, preshift_gstride, x0_lo, x1_lo);
  B200_CHECK(false, "groupnorm: bad act_dtype");
}

extern "C" int b200_add4_scale_stats(const void *a, const void *b, const void *c, const void *d, float scale, void *dst,
                                     int N, int HW, int C, double *sums, const void *a_lo, const void *b_lo,
                                     const void *c_lo, const void *d_lo, void *dst_lo, int act_dtype, void *stream) {
  using namespace b200;
  B200_CHECK(a && dst && sums && N > 0 && HW > 0 && C > 0 && C % 8 == 0 && C <= 8 * kNormThreads,
             "add4_scale_stats: bad arguments");
  dim3 grid;
  int ppb;
  stats_grid(N, HW, C >> 3, &grid, &ppb);
  if (!b) b_lo = nullptr;
  if (!c) c_lo = nullptr;
  if (!d) d_lo = nullptr;
  if (a_lo || b_lo || c_lo || d_lo || dst_lo) {
    if (act_dtype == B200_F16)
      add4_stats_lo_kernel<__half><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>(
          (const __half *)a, (const __half *)b, (const __half *)c, (const __half *)d, scale, (__half *)dst, C, HW, ppb, sums,
          (const __half *)a_lo, (const __half *)b_lo, (const __half *)c_lo, (const __half *)d_lo, (__half *)dst_lo);
    else if (act_dtype == B200_BF16)
      add4_stats_lo_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16 *)a, (const __nv_bfloat16 *)b, (const __nv_bfloat16 *)c, (const __nv_bfloat16 *)d, scale,
          (__nv_bfloat16 *)dst, C, HW, ppb, sums, (const __nv_bfloat16 *)a_lo, (const __nv_bfloat16 *)b_lo,
          (const __nv_bfloat16 *)c_lo, (const __nv_bfloat16 *)d_lo, (__nv_bfloat16 *)dst_lo);
    else
      B200_CHECK(false, "add4_scale_stats: bad act_dtype");
    B200_LAUNCH_CHECK();
    return 0;
  }
  if (act_dtype == B200_F16)
    add4_stats_kernel<__half><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>(
        (const __half *)a, (const __half *)b, (const __half *)c, (const __half *)d, scale, (__half *)dst, C, HW, ppb, sums);
  else if (act_dtype == B200_BF16)
    add4_stats_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)a, (const __nv_bfloat16 *)b, (const __nv_bfloat16 *)c, (const __nv_bfloat16 *)d, scale,
        (__nv_bfloat16 *)dst, C, HW, ppb, sums);
  else
    B200_CHECK(false, "add4_scale_stats: bad act_dtype");
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_channel_stats(const void *x, int N, int HW, int C, double *sums, int act_dtype, void *stream) {
  using namespace b200;
  B200_CHECK(x && sums && N > 0 && HW > 0 && C > 0 && C % 8 == 0 && C <= 8 * kNormThreads, "channel_stats: bad arguments");
  dim3 grid;
  int ppb;
  stats_grid(N, HW, C >> 3, &grid, &ppb);
  if (act_dtype == B200_F16)
    channel_stats_kernel<__half><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>((const __half *)x, C, HW, ppb, sums);
  else if (act_dtype == B200_BF16)
    channel_stats_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, C, HW, ppb, sums);
  else
    B200_CHECK(false, "channel_stats: bad act_dtype");
  B200_LAUNCH_CHECK();
  return 0;
}
