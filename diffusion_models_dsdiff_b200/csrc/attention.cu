// Fused attention for sm_100a: softmax(scale * Q K^T) V without materialising the Lq x Lk matrix.
//
// Replaces QKVAttention / QKVAttentionLegacy (ldm/modules/diffusionmodules/openaimodel.py:537-555 / 505-521),
// CrossAttention's einsum/softmax/einsum (ldm/modules/attention.py:163-194) and timm Attention as used by DiT
// (UNet_DS_Diff/DiT_models.py:108).
//
// flash_attn_kernel (tcgen05): one CTA = 128 queries of one (batch, head).
//   * both GEMMs run on the tensor cores: S = Q K^T (M=128 queries, N=128 keys, K=d) and O += P V (M=128, N=d,
//     K=128 keys), fp32 accumulators in TMEM (S double-buffered, O resident for the whole key loop);
//   * operands are staged by TMA as un-swizzled 8x16-byte core matrices: one box of (8 channels x 128 tokens) per
//     16-byte channel block, so Q and K land K-major and V lands MN-major with NO transpose pass;
//   * ONE sweep over the keys with an online softmax: per key tile the row maximum of the tile (two warps per row
//     exchange their 64-column halves through shared memory + a 64-thread named barrier) updates a running reference M
//     LAZILY — only when it grows by more than 2^8 — so P = exp2((S - M) * scale*log2e) stays <= 256 (exact in 16-bit)
//     and the O accumulator in TMEM is rescaled (tcgen05.ld -> mul -> tcgen05.st by the softmax warps, after the
//     previous P V has completed) only on those rare steps.  S is read from TMEM once per tile: the kernel is bound by
//     the TMEM read port (64 B/clk) and the exponentials (MUFU), both of which the old exact two-sweep version paid twice.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 = softmax: thread == (query row,
//     64-column half of the key tile), two warps per scheduler so MUFU / TMEM-load latency overlaps.
//
// attn_simt_kernel: the v1 fp32 CUDA-core kernel, kept for head dims that are not a multiple of 16.
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// tcgen05 flash attention
// ------------------------------------------------------------------------------------------
constexpr int kFaThreads = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (4 lane quadrants x 2 column halves)
constexpr int kFaBM = 128;           // queries per CTA
constexpr int kFaBN = 128;           // keys per tile
constexpr int kFaSCols = 128;        // TMEM columns per S buffer
constexpr int kFaPBytes = kFaBM * kFaBN * 2;
constexpr uint32_t kCoreRowBytes = 16;
constexpr uint32_t kBlkBytes = 128 * kCoreRowBytes;   // one (8 channels x 128 tokens) TMA box = 2048 B

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 8 probabilities in [0, 1] -> 16-bit (no saturation needed)
template <typename T>
__device__ __forceinline__ uint4 pack8_unit(const float (&f)[8]);
template <>
__device__ __forceinline__ uint4 pack8_unit<__half>(const float (&f)[8]) {
  uint4 u;
  __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
template <>
__device__ __forceinline__ uint4 pack8_unit<__nv_bfloat16>(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

struct FaMaps {
  CUtensorMap q, k, v;
};

struct FaParams {
  int Lq, Lk, heads, B;
  int q_off, k_off, v_off;            // first channel of head 0 inside each tensor's rows
  int q_hstride, k_hstride, v_hstride;  // channels between consecutive heads
  int ldo;                            // output row pitch (elements); head h writes channels [h*D, h*D+D)
  int stages;                         // K / V ring depth (1 or 2)
  int pbufs;                          // P buffers in shared memory (2, or 1 to fit two CTAs per SM at D = 64)
  float scale_log2e;
  void *out;
  float *lse;                         // optional [B, heads, Lq]: log-sum-exp of the scaled scores (training)
  uint32_t idesc_s, idesc_o;
};

template <typename T, int D>
__global__ void __launch_bounds__(kFaThreads, (D <= 64) ? 2 : 1)
flash_attn_kernel(const __grid_constant__ FaMaps maps, const FaParams p) {
  constexpr int kDB = D / 8;                         // 16-byte channel blocks per row
  constexpr uint32_t kTileBytes = kDB * kBlkBytes;   // Q / K / V tile: D x 128 x 2 B
  // D == 64: a token row is one 128-byte swizzle row, so every Q / K / V tile is ONE SWIZZLE_128B TMA box (128 row fetches of
  // 128 B instead of 1024 of 16 B) whose shared-memory image is the canonical K-major tile for Q K^T and the canonical
  // MN-major tile for P V (see attention_bwd.cu)
  constexpr bool kSw = (D == 64);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, o_full;
  __shared__ __align__(8) uint64_t k_full[2], k_empty[2], v_full[2], v_empty[2];
  __shared__ __align__(8) uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2];
  __shared__ __align__(8) uint64_t pv_done;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_xchg[2][2][kFaBM];  // [tile parity][column half][row]: per-row max / sum exchange between the halves

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kFaBM;
  const int h = blockIdx.y, b = blockIdx.z;
  const int nk = (p.Lk + kFaBN - 1) / kFaBN;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B tiles: 1024-byte alignment
  const uint32_t sQ = base;
  const uint32_t sK = sQ + kTileBytes;
  const uint32_t sV = sK + p.stages * kTileBytes;
  const uint32_t sP = sV + p.stages * kTileBytes;

  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.q); prefetch_tmap(&maps.k); prefetch_tmap(&maps.v);
    mbar_init(smem_u32(&q_full), 1);
    mbar_init(smem_u32(&o_full), 1);
    mbar_init(smem_u32(&pv_done), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&k_full[i]), 1);  mbar_init(smem_u32(&k_empty[i]), 1);
      mbar_init(smem_u32(&v_full[i]), 1);  mbar_init(smem_u32(&v_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);  mbar_init(smem_u32(&s_empty[i]), 8);
      mbar_init(smem_u32(&p_full[i]), 8);  mbar_init(smem_u32(&p_empty[i]), 1);
    }
    fence_barrier_init();
  }
  // S (128 columns, single buffer: the softmax warps copy it to registers and hand it back at once) + O (D columns)
  constexpr uint32_t kTmemCols = (D <= 128) ? 256 : 512;
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_o = tmem_base + kFaSCols;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t qb = smem_u32(&q_full);
      mbar_arrive_expect_tx(qb, kTileBytes);
      for (int c = 0; c < (kSw ? 1 : kDB); ++c)
        tma_load_3d(sQ + c * kBlkBytes, &maps.q, qb, p.q_off + h * p.q_hstride + c * 8, q0, b);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < nk; ++j) {
        mbar_wait(smem_u32(&k_empty[ks]), kph ^ 1);
        const uint32_t kb = smem_u32(&k_full[ks]);
        mbar_arrive_expect_tx(kb, kTileBytes);
        for (int c = 0; c < (kSw ? 1 : kDB); ++c)
          tma_load_3d(sK + ks * kTileBytes + c * kBlkBytes, &maps.k, kb, p.k_off + h * p.k_hstride + c * 8,
                      j * kFaBN, b);
        if (++ks == p.stages) { ks = 0; kph ^= 1; }
        mbar_wait(smem_u32(&v_empty[vs]), vph ^ 1);
        const uint32_t vb = smem_u32(&v_full[vs]);
        mbar_arrive_expect_tx(vb, kTileBytes);
        for (int c = 0; c < (kSw ? 1 : kDB); ++c)
          tma_load_3d(sV + vs * kTileBytes + c * kBlkBytes, &maps.v, vb, p.v_off + h * p.v_hstride + c * 8,
                      j * kFaBN, b);
        if (++vs == p.stages) { vs = 0; vph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_wait(smem_u32(&q_full), 0);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      int g = 0;                       // global S-tile counter (both sweeps)
      auto issue_qk = [&](int gi) {
        const int sb = 0;
        mbar_wait(smem_u32(&k_full[ks]), kph);
        mbar_wait(smem_u32(&s_empty[sb]), (gi & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          // 16 channels = 2 core-matrix columns: advance 2 * LBO
          const uint64_t da = kSw ? umma_desc_sw128(sQ) + (uint64_t)(kk * 2) : umma_desc_nosw(sQ + kk * 2 * kBlkBytes, kBlkBytes, 128);
          const uint64_t db = kSw ? umma_desc_sw128(sK + ks * kTileBytes) + (uint64_t)(kk * 2)
                                  : umma_desc_nosw(sK + ks * kTileBytes + kk * 2 * kBlkBytes, kBlkBytes, 128);
          umma_f16(tmem_base + sb * kFaSCols, da, db, p.idesc_s, kk > 0);
        }
        umma_commit(smem_u32(&k_empty[ks]));
        umma_commit(smem_u32(&s_full[sb]));
        if (++ks == p.stages) { ks = 0; kph ^= 1; }
      };
      // scores run one tile ahead of P V
      issue_qk(g++);
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) issue_qk(g++);
        const int pb = (p.pbufs == 2) ? (j & 1) : 0;
        mbar_wait(smem_u32(&v_full[vs]), vph);
        mbar_wait(smem_u32(&p_full[pb]), (p.pbufs == 2) ? ((j >> 1) & 1) : (j & 1));
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kFaBN / 16; ++kk) {
          // A = P, K-major [key block][row][8]: 16 keys = 2 blocks of 2048 B.
          const uint64_t da = umma_desc_nosw(sP + pb * kFaPBytes + kk * 2 * kBlkBytes, kBlkBytes, 128);
          // B = V, MN-major [channel block][key][8 ch]: along K (keys) core matrices are 128 B apart,
          // along N (channels) 2048 B apart.
          const uint64_t db = kSw ? umma_desc_sw128(sV + vs * kTileBytes + kk * 2048)      // 8-key atoms 1024 B apart
                                  : umma_desc_nosw(sV + vs * kTileBytes + kk * 2 * 128, 128, kBlkBytes);
          umma_f16(tmem_o, da, db, p.idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&v_empty[vs]));
        umma_commit(smem_u32(&p_empty[pb]));
        umma_commit(smem_u32(&pv_done));       // O holds tiles 0..j: the softmax warps may rescale it
        if (++vs == p.stages) { vs = 0; vph ^= 1; }
      }
      umma_commit(smem_u32(&o_full));
    }
    __syncwarp();
  } else {
    // ===================== softmax warps: thread == (query row, column half) =====================
    const int lane_base = 32 * (warp & 3);
    const int row = lane_base + lane;
    const int half = (warp - 2) >> 2;                 // key columns [64*half, 64*half + 64) of every tile
    const int cbase = half * 64;
    const uint32_t t_lane = (uint32_t)lane_base << 16;
    const int pair_bar = 2 + (warp & 3);              // named barrier of the two warps that share this row quadrant
    constexpr int kHalfD = D / 2;                     // O columns this half rescales / writes out (multiple of 8)
    float M = -INFINITY;                              // running softmax reference (raw-score units), lazily updated
    float l = 0.f;
    for (int j = 0; j < nk; ++j) {
      const int sb = 0, pb = (p.pbufs == 2) ? (j & 1) : 0;
      mbar_wait(smem_u32(&s_full[sb]), j & 1);
      tc_fence_after();
      const int nvalid = min(kFaBN, p.Lk - j * kFaBN);
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_base + t_lane + sb * kFaSCols + cbase, r0);
      tmem_ld32(tmem_base + t_lane + sb * kFaSCols + cbase + 32, r1);
      tmem_ld_wait();
      // S tile is in registers: hand the TMEM buffer back so the next Q K^T can start
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[sb]));
      const bool full = nvalid == kFaBN;                // every tile but (possibly) the last: no column masks
      float mt = -INFINITY;
      if (full) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};       // 4 independent chains
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          m4[i & 3] = fmaxf(m4[i & 3], fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
        }
        mt = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (cbase + i < nvalid) mt = fmaxf(mt, __uint_as_float(r0[i]));
          if (cbase + 32 + i < nvalid) mt = fmaxf(mt, __uint_as_float(r1[i]));
        }
      }
      s_xchg[j & 1][half][row] = mt;
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      mt = fmaxf(mt, s_xchg[j & 1][half ^ 1][row]);
      // lazy reference update: both halves of a row take the same decision from the same numbers
      float alpha = 1.f;
      bool grow = false;
      if (j == 0) {
        M = mt;
      } else if ((mt - M) * p.scale_log2e > 8.f) {
        alpha = ex2_approx((M - mt) * p.scale_log2e);
        M = mt;
        l *= alpha;
        grow = true;
      }
      if (__any_sync(0xffffffffu, grow)) {
        // rescale this row's half of the O accumulator; P V of tile j-1 must have landed first
        mbar_wait(smem_u32(&pv_done), (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < kHalfD; c0 += 8) {
          uint32_t o8[8];
          tmem_ld8(tmem_o + t_lane + half * kHalfD + c0, o8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * alpha);
          tmem_st8(tmem_o + t_lane + half * kHalfD + c0, o8);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      const float mneg = -M * p.scale_log2e;
      mbar_wait(smem_u32(&p_empty[pb]), ((p.pbufs == 2) ? ((j >> 1) & 1) : (j & 1)) ^ 1);
      uint8_t *prow = smem_raw + (sP - smem_u32(smem_raw)) + pb * kFaPBytes + row * 16;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        const uint32_t (&r)[32] = c0 ? r1 : r0;
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          float f[8];
          if (full) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = ex2_approx(fmaf(__uint_as_float(r[v8 * 8 + i]), p.scale_log2e, mneg));
            l += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = cbase + c0 + v8 * 8 + i;
              const float e = ex2_approx(fmaf(__uint_as_float(r[v8 * 8 + i]), p.scale_log2e, mneg));
              f[i] = c < nvalid ? e : 0.f;
              l += f[i];
            }
          }
          const uint4 u = pack8_sat<T>(f);
          *reinterpret_cast<uint4 *>(prow + (((cbase + c0) >> 3) + v8) * kBlkBytes) = u;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[pb]));
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");      // every pair is past its last per-tile exchange
    s_xchg[0][half][row] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l = s_xchg[0][0][row] + s_xchg[0][1][row];
    // epilogue: O / l -> out[b, q0+row, h*D ...]; the halves split the D output channels
    mbar_wait(smem_u32(&o_full), 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int qi = q0 + row;
    if (p.lse && half == 0 && qi < p.Lq)      // P = exp2((S - M) scale log2e) -> lse = M scale + ln(l)   (natural-log units)
      p.lse[((size_t)b * p.heads + h) * p.Lq + qi] = M * p.scale_log2e * 0.6931471805599453f + logf(l);
    T *o = reinterpret_cast<T *>(p.out) + ((size_t)b * p.Lq + qi) * p.ldo + h * D;
#pragma unroll
    for (int c0 = 0; c0 < kHalfD; c0 += 8) {
      uint32_t r[8];
      tmem_ld8(tmem_o + t_lane + half * kHalfD + c0, r);
      tmem_ld_wait();
      if (qi < p.Lq) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[i]) * inv;
        *reinterpret_cast<uint4 *>(o + half * kHalfD + c0) = pack8<T>(f);
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

struct AttnArgs {
  const void *q, *k, *v;
  void *out;
  float *lse;
  int B, Lq, Lk, heads, d;
  int ldq, ldk, ldv, ldo;
  int q_off, k_off, v_off, q_hstride, k_hstride, v_hstride;
  float scale;
};

template <typename T, int D>
static int launch_flash(const AttnArgs &a, cudaStream_t st) {
  FaMaps maps{};
  FaParams p{};
  const void *ptr[3] = {a.q, a.k, a.v};
  const int ld[3] = {a.ldq, a.ldk, a.ldv};
  const int len[3] = {a.Lq, a.Lk, a.Lk};
  CUtensorMap *mp[3] = {&maps.q, &maps.k, &maps.v};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[3] = {(uint64_t)ld[i], (uint64_t)len[i], (uint64_t)a.B};
    uint64_t str[2] = {(uint64_t)ld[i] * 2, (uint64_t)len[i] * ld[i] * 2};
    uint32_t box[3] = {(uint32_t)(D == 64 ? 64 : 8), 128, 1};
    uint32_t es[3] = {1, 1, 1};
    if (int rc = cached_tmap(mp[i], ptr[i], 3, dims, str, box, es, /*swizzle128=*/D == 64)) return rc;
  }
  p.Lq = a.Lq; p.Lk = a.Lk; p.heads = a.heads; p.B = a.B;
  p.q_off = a.q_off; p.k_off = a.k_off; p.v_off = a.v_off;
  p.q_hstride = a.q_hstride; p.k_hstride = a.k_hstride; p.v_hstride = a.v_hstride;
  p.ldo = a.ldo;
  p.out = a.out;
  p.lse = a.lse;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  p.idesc_s = umma_idesc(Act<T>::kFmt, kFaBM, kFaBN, 0);
  p.idesc_o = umma_idesc(Act<T>::kFmt, kFaBM, D, 1);
  const size_t tile = (size_t)D * 256;
  // D <= 64: two CTAs per SM (256 TMEM columns each), so one CTA's TMEM-read phase overlaps the other's exponentials;
  // its shared memory must stay under ~110 KB
  const size_t budget = (D <= 64) ? 108 * 1024 : 200 * 1024;
  p.pbufs = (tile * 3 + 2 * kFaPBytes + 256 <= budget) ? 2 : 1;      // D = 64: one P buffer so that two CTAs fit
  p.stages = (tile * 5 + p.pbufs * kFaPBytes + 256 <= budget) ? 2 : 1;
  const size_t smem = tile * (1 + 2 * p.stages) + p.pbufs * kFaPBytes + 1024;
  B200_CHECK(smem <= 220 * 1024, "attention: head dim %d does not fit shared memory", D);
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(flash_attn_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    B200_CUDA(cudaFuncSetAttribute(flash_attn_kernel<T, D>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  const dim3 grid(cdiv(a.Lq, kFaBM), a.heads, a.B);
  flash_attn_kernel<T, D><<<grid, kFaThreads, smem, st>>>(maps, p);
  B200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// v1 CUDA-core kernel (head dims that are not a multiple of 16)
// ------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 128;
constexpr int kKT = 32;  // keys per tile

template <typename T, int CH>
__global__ void __launch_bounds__(kAttnThreads)
attn_simt_kernel(const AttnArgs a, float scale_log2e) {
  __shared__ __align__(16) float sK[kKT][CH];
  __shared__ __align__(16) float sV[kKT][CH];
  const int h = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * kAttnThreads + threadIdx.x;
  const bool valid = qi < a.Lq;
  const T *qb = reinterpret_cast<const T *>(a.q) + (size_t)b * a.Lq * a.ldq + a.q_off + h * a.q_hstride;
  const T *kb = reinterpret_cast<const T *>(a.k) + (size_t)b * a.Lk * a.ldk + a.k_off + h * a.k_hstride;
  const T *vb = reinterpret_cast<const T *>(a.v) + (size_t)b * a.Lk * a.ldv + a.v_off + h * a.v_hstride;

  float q[CH], acc[CH];
#pragma unroll
  for (int c = 0; c < CH; c += 8) {
    float f[8];
    if (valid) unpack8<T>(*reinterpret_cast<const uint4 *>(qb + (size_t)qi * a.ldq + c), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      q[c + i] = valid ? f[i] * scale_log2e : 0.f;
      acc[c + i] = 0.f;
    }
  }
  float m = -INFINITY, l = 0.f;

  for (int kt = 0; kt < a.Lk; kt += kKT) {
    __syncthreads();
    constexpr int kVecPerRow = CH / 8;
    for (int i = threadIdx.x; i < kKT * kVecPerRow * 2; i += kAttnThreads) {
      const int which = i / (kKT * kVecPerRow);
      const int r = (i % (kKT * kVecPerRow)) / kVecPerRow;
      const int cv = (i % kVecPerRow) * 8;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (kt + r < a.Lk)
        unpack8<T>(*reinterpret_cast<const uint4 *>((which ? vb + (size_t)(kt + r) * a.ldv : kb + (size_t)(kt + r) * a.ldk) + cv), f);
      float *dst = which ? &sV[r][cv] : &sK[r][cv];
      *reinterpret_cast<float4 *>(dst) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4 *>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    const int nkeys = min(kKT, a.Lk - kt);
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
      s[j] = j < nkeys ? d : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = exp2f(m - m_new);   // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] *= corr;
#pragma unroll
    for (int j = 0; j < kKT; ++j) {
      const float pj = exp2f(s[j] - m_new);
      l += pj;
#pragma unroll
      for (int c = 0; c < CH; c += 4) {
        const float4 vv = *reinterpret_cast<const float4 *>(&sV[j][c]);
        acc[c] = fmaf(pj, vv.x, acc[c]); acc[c + 1] = fmaf(pj, vv.y, acc[c + 1]);
        acc[c + 2] = fmaf(pj, vv.z, acc[c + 2]); acc[c + 3] = fmaf(pj, vv.w, acc[c + 3]);
      }
    }
    m = m_new;
  }
  if (valid) {
    const float inv = 1.f / l;
    T *o = reinterpret_cast<T *>(a.out) + ((size_t)b * a.Lq + qi) * a.ldo + h * CH;
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
static int launch_attn(const AttnArgs &a, int force_simt, cudaStream_t st) {
  if (!force_simt) {
    switch (a.d) {
      case 32: return launch_flash<T, 32>(a, st);
      case 48: return launch_flash<T, 48>(a, st);
      case 64: return launch_flash<T, 64>(a, st);
      case 80: return launch_flash<T, 80>(a, st);
      case 96: return launch_flash<T, 96>(a, st);
      case 128: return launch_flash<T, 128>(a, st);
      case 160: return launch_flash<T, 160>(a, st);
      default: break;
    }
  }
  const dim3 grid(cdiv(a.Lq, kAttnThreads), a.heads, a.B);
  const float sl = a.scale * 1.4426950408889634f;
#define ATTN_CASE(CHV)                                                        \
  case CHV:                                                                   \
    attn_simt_kernel<T, CHV><<<grid, kAttnThreads, 0, st>>>(a, sl);           \
    break;
  switch (a.d) {
    ATTN_CASE(8) ATTN_CASE(16) ATTN_CASE(24) ATTN_CASE(32) ATTN_CASE(40) ATTN_CASE(48) ATTN_CASE(64)
    default: B200_CHECK(false, "attention: unsupported head dim %d", a.d);
  }
#undef ATTN_CASE
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200

extern "C" int b200_attention(const b200_attn_desc *d, void *stream) {
  using namespace b200;
  B200_CHECK(d != nullptr && d->q && d->k && d->v && d->out, "attention: null pointer");
  B200_CHECK(d->B > 0 && d->Lq > 0 && d->Lk > 0 && d->heads > 0 && d->d > 0 && d->d % 8 == 0,
             "attention: bad shape (head dim must be a multiple of 8)");
  B200_CHECK(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0 && d->q_off % 8 == 0 &&
                 d->k_off % 8 == 0 && d->v_off % 8 == 0 && d->q_hstride % 8 == 0 && d->k_hstride % 8 == 0 &&
                 d->v_hstride % 8 == 0,
             "attention: pitches and offsets must be multiples of 8 elements (16 bytes)");
  B200_CHECK((((uintptr_t)d->q | (uintptr_t)d->k | (uintptr_t)d->v | (uintptr_t)d->out) & 15) == 0,
             "attention: unaligned pointer");
  AttnArgs a;
  a.q = d->q; a.k = d->k; a.v = d->v; a.out = d->out;
  a.lse = d->lse;
  B200_CHECK(d->lse == nullptr || (d->d % 16 == 0 && !d->force_simt), "attention: lse needs the tcgen05 kernel (d %% 16 == 0)");
  a.B = d->B; a.Lq = d->Lq; a.Lk = d->Lk; a.heads = d->heads; a.d = d->d;
  a.ldq = d->ldq; a.ldk = d->ldk; a.ldv = d->ldv; a.ldo = d->ldo;
  a.q_off = d->q_off; a.k_off = d->k_off; a.v_off = d->v_off;
  a.q_hstride = d->q_hstride; a.k_hstride = d->k_hstride; a.v_hstride = d->v_hstride;
  a.scale = d->scale;
  if (d->act_dtype == B200_F16) return launch_attn<__half>(a, d->force_simt, (cudaStream_t)stream);
  if (d->act_dtype == B200_BF16) return launch_attn<__nv_bfloat16>(a, d->force_simt, (cudaStream_t)stream);
  B200_CHECK(false, "attention: bad act_dtype");
}

extern "C" int b200_attention_qkv(const void *qkv, void *out, int B, int L, int heads, int ch, int new_order,
                                  float scale, int act_dtype, void *stream) {
  b200_attn_desc d;
  const int C = heads * ch;
  d.q = d.k = d.v = qkv;
  d.out = out;
  d.B = B; d.Lq = L; d.Lk = L; d.heads = heads; d.d = ch;
  d.ldq = d.ldk = d.ldv = 3 * C;
  d.ldo = C;
  if (new_order) {   // q | k | v blocks of heads*ch (openaimodel.py:537-555)
    d.q_off = 0; d.k_off = C; d.v_off = 2 * C;
    d.q_hstride = d.k_hstride = d.v_hstride = ch;
  } else {           // per head [q, k, v] (QKVAttentionLegacy, :505-521)
    d.q_off = 0; d.k_off = ch; d.v_off = 2 * ch;
    d.q_hstride = d.k_hstride = d.v_hstride = 3 * ch;
  }
  d.scale = scale;
  d.act_dtype = act_dtype;
  d.force_simt = 0;
  d.lse = nullptr;
  return b200_attention(&d, stream);
}
