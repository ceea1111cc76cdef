// Attention BACKWARD on tcgen05 / TMEM (training step of DiT): the gradient of
//   O = softmax(scale Q K^T) V      (timm Attention inside DiTBlock, UNet_DS_Diff/DiT_models.py:108)
// which the reference obtains from torch autograd (loss.backward(), trainers/trainer_use_gaussian_diff_adv.py:345-382).
//
//   S = Q K^T,  P = exp(scale S - lse),  dP = dO V^T,  dS = scale P (dP - D),  D_i = sum_c dO[i, c] O[i, c]
//   dV = P^T dO,   dK = dS^T Q,   dQ = dS K
//
// Two instantiations of ONE kernel, so that no accumulator is shared between CTAs (no atomics, deterministic):
//   kDQ = false : CTA = 128 KEYS of one (batch, head); streams the 128-query tiles; dK / dV accumulate in TMEM;
//   kDQ = true  : CTA = 128 QUERIES; streams the 128-key tiles; dQ accumulates in TMEM.
// Both recompute S and dP per tile pair (7 GEMMs per pair instead of the minimal 5).  Same machinery as the forward kernel
// (attention.cu): operands staged by TMA as UN-SWIZZLED 8 x 16-byte core matrices, one box of (8 channels x 128 tokens)
// per 16-byte channel block — that one shared-memory image serves as a K-major operand when the contraction runs over
// channels (S, dP) and as an MN-major operand when it runs over tokens (dV, dK, dQ), so nothing is ever transposed;
// warp 0 = TMA producer, warp 1 = single-thread MMA issuer + TMEM owner, warps 2-17 = the element-wise stage
// (thread = (query row, 32-column quarter)): tcgen05.ld S and dP -> registers, hand the TMEM buffers back at once so the
// next tile's S / dP MMAs overlap the exponentials; P and dS are computed and packed in registers BEFORE waiting for the
// previous tile's dV / dK / dQ MMAs to release the operand tiles, so those MMAs overlap this tile's arithmetic too.
// TMEM columns: S 128 | dP 128 | accumulator 0 (D) | accumulator 1 (D).
#include "common.cuh"

namespace b200 {

constexpr int kBwThreads = 576;                          // warp 0 TMA, warp 1 MMA, warps 2-17 element-wise
constexpr int kBwT = 128;                                // tokens per tile (both queries and keys)
constexpr uint32_t kBwBlk = 128 * 16;                    // one (8 channels x 128 tokens) TMA box = 2048 B
constexpr uint32_t kBwPBytes = kBwT * kBwT * 2;          // a 128 x 128 16-bit operand tile (P or dS)

struct BwMaps {
  CUtensorMap qkv, dout;        // 128-token boxes
  CUtensorMap qkv_s, dout_s;    // 64-token boxes (v2: streamed tiles)
};

struct BwParams {
  int L, heads;
  int q_off, k_off, v_off;      // first channel of head 0 inside a qkv row
  int ld;                       // qkv / dqkv row pitch (elements)
  float scale, scale_log2e;
  const float *lse;             // [B, heads, L] natural-log units (b200_attention)
  const float *dsum;            // [B, L, heads]
  void *dqkv;
  uint32_t idesc_s, idesc_acc;
  int debug;                    // B200DIFF_DEBUG_TOGGLES builds only: timing experiments (results are garbage)
};

__device__ __forceinline__ float bw_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, int D, bool kDQ>
__global__ void __launch_bounds__(kBwThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ BwMaps maps, const BwParams p) {
  constexpr int kDB = D / 8;
  constexpr uint32_t kTile = kDB * kBwBlk;               // a [128 tokens x D] operand tile
  // D == 64: a token row is exactly one 128-byte swizzle row, so each Q / K / V / dO tile is ONE TMA box of (64 channels x
  // 128 tokens) with SWIZZLE_128B — 128 row fetches of 128 B instead of 1024 fetches of 16 B.  The same shared-memory
  // image is the canonical K-major SW128 tile (rows = tokens = M/N, 128 B = 64 K elements) AND the canonical MN-major SW128
  // tile (rows = tokens = K, 128 B = 64 MN elements; 8-row atoms 1024 B apart along K).
  constexpr bool kSw = (D == 64);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t res_full, acc_full, s_full, s_empty, p_full, p_empty;
  constexpr int NS = kDQ ? 4 : 3;                        // stages of the streamed pair (TMA latency is the loop's critical path)
  __shared__ __align__(8) uint64_t str_full[NS], str_empty[NS];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * kBwT;                      // first resident token (keys when !kDQ, queries when kDQ)
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_it = (p.L + kBwT - 1) / kBwT;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t sRes0 = base, sRes1 = base + kTile;     // resident pair: (K, V) or (Q, dO)
  const uint32_t sStr = sRes1 + kTile;                   // streamed pair x 2 stages: (Q, dO) or (K, V)
  const uint32_t sdS = sStr + 2 * NS * kTile;
  const uint32_t sP = sdS + kBwPBytes;                   // only when !kDQ

  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.qkv); prefetch_tmap(&maps.dout);
    mbar_init(smem_u32(&res_full), 1);
    mbar_init(smem_u32(&acc_full), 1);
    mbar_init(smem_u32(&s_full), 1);
    mbar_init(smem_u32(&s_empty), 16);
    mbar_init(smem_u32(&p_full), 16);
    mbar_init(smem_u32(&p_empty), 1);
    for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&str_full[i]), 1); mbar_init(smem_u32(&str_empty[i]), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128;
  const uint32_t tmem_a0 = tmem_base + 256, tmem_a1 = tmem_base + 256 + D;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int qc = p.q_off + h * D, kc = p.k_off + h * D, vc = p.v_off + h * D, oc = h * D;
      {
        const uint32_t bar = smem_u32(&res_full);
        mbar_arrive_expect_tx(bar, 2 * kTile);
        for (int c = 0; c < (kSw ? 1 : kDB); ++c) {
          if (kDQ) {
            tma_load_3d(sRes0 + c * kBwBlk, &maps.qkv, bar, qc + c * 8, t0, b);
            tma_load_3d(sRes1 + c * kBwBlk, &maps.dout, bar, oc + c * 8, t0, b);
          } else {
            tma_load_3d(sRes0 + c * kBwBlk, &maps.qkv, bar, kc + c * 8, t0, b);
            tma_load_3d(sRes1 + c * kBwBlk, &maps.qkv, bar, vc + c * 8, t0, b);
          }
        }
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it % NS;
        mbar_wait(smem_u32(&str_empty[st]), ((it / NS) & 1) ^ 1);
        const uint32_t bar = smem_u32(&str_full[st]);
        mbar_arrive_expect_tx(bar, 2 * kTile);
        const uint32_t d0 = sStr + st * 2 * kTile, d1 = d0 + kTile;
        for (int c = 0; c < (kSw ? 1 : kDB); ++c) {
          if (kDQ) {
            tma_load_3d(d0 + c * kBwBlk, &maps.qkv, bar, kc + c * 8, it * kBwT, b);
            tma_load_3d(d1 + c * kBwBlk, &maps.qkv, bar, vc + c * 8, it * kBwT, b);
          } else {
            tma_load_3d(d0 + c * kBwBlk, &maps.qkv, bar, qc + c * 8, it * kBwT, b);
            tma_load_3d(d1 + c * kBwBlk, &maps.dout, bar, oc + c * 8, it * kBwT, b);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_wait(smem_u32(&res_full), 0);
      // S = Q K^T and dP = dO V^T of tile `g`: contraction over the D channels, every operand K-major
      auto issue_first = [&](int g) {
        const int st = g % NS;
        mbar_wait(smem_u32(&str_full[st]), (g / NS) & 1);
        mbar_wait(smem_u32(&s_empty), (g & 1) ^ 1);
        tc_fence_after();
        const uint32_t x0 = sStr + st * 2 * kTile, x1 = x0 + kTile;
        const uint32_t aQ = kDQ ? sRes0 : x0, bK = kDQ ? x0 : sRes0;
        const uint32_t aO = kDQ ? sRes1 : x1, bV = kDQ ? x1 : sRes1;
        // 16 channels per instruction: two un-swizzled core-matrix columns (+2 boxes), or +32 B inside the swizzled row
        auto kdesc = [&](uint32_t tile, int kk) -> uint64_t {
          return kSw ? umma_desc_sw128(tile) + (uint64_t)(kk * 2) : umma_desc_nosw(tile + kk * 2 * kBwBlk, kBwBlk, 128);
        };
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) umma_f16(tmem_s, kdesc(aQ, kk), kdesc(bK, kk), p.idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) umma_f16(tmem_dp, kdesc(aO, kk), kdesc(bV, kk), p.idesc_s, kk > 0);
        umma_commit(smem_u32(&s_full));
      };
      issue_first(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_first(it + 1);          // runs while the element-wise warps work on tile `it`
        const int st = it % NS;
        mbar_wait(smem_u32(&p_full), it & 1);
        tc_fence_after();
        const uint32_t x0 = sStr + st * 2 * kTile, x1 = x0 + kTile;
        const uint32_t acc = (it > 0) ? 1u : 0u;
        // B operand read MN-major over 16 tokens: un-swizzled core matrices 128 B apart along K (tokens), 2048 B along N
        // (+256 B per 16 tokens); swizzled: 8-token atoms 1024 B apart (+2048 B per 16 tokens)
        auto mndesc = [&](uint32_t tile, int kk) -> uint64_t {
          return kSw ? umma_desc_sw128(tile + kk * 2048) : umma_desc_nosw(tile + kk * 256, 128, kBwBlk);
        };
#ifdef B200DIFF_DEBUG_TOGGLES
        if (p.debug & 8) {
          umma_commit(smem_u32(&str_empty[st]));
          umma_commit(smem_u32(&p_empty));
          continue;
        }
#endif
        if (!kDQ) {
          // dV += P^T dO, dK += dS^T Q: contraction over the 128 QUERIES.  A = P / dS tile [key block][query][8 keys] read
          // MN-major (8 keys contiguous, queries 16 B apart: core matrices 128 B apart along K, 2048 B along M);
          // B = dO / Q tile [channel block][query][8 ch] read MN-major.  16 queries per instruction = +256 B.
#pragma unroll
          for (int kk = 0; kk < kBwT / 16; ++kk)
            umma_f16(tmem_a0, umma_desc_nosw(sP + kk * 256, 128, kBwBlk), mndesc(x1, kk), p.idesc_acc, acc | (uint32_t)(kk > 0));
#pragma unroll
          for (int kk = 0; kk < kBwT / 16; ++kk)
            umma_f16(tmem_a1, umma_desc_nosw(sdS + kk * 256, 128, kBwBlk), mndesc(x0, kk), p.idesc_acc, acc | (uint32_t)(kk > 0));
        } else {
          // dQ += dS K: contraction over the 128 KEYS.  A = dS tile read K-major (16 keys = 2 key blocks = +4096 B);
          // B = K tile [channel block][key][8 ch] read MN-major (+256 B per 16 keys).
#pragma unroll
          for (int kk = 0; kk < kBwT / 16; ++kk)
            umma_f16(tmem_a0, umma_desc_nosw(sdS + kk * 2 * kBwBlk, kBwBlk, 128), mndesc(x0, kk), p.idesc_acc,
                     acc | (uint32_t)(kk > 0));
        }
        umma_commit(smem_u32(&str_empty[st]));
        umma_commit(smem_u32(&p_empty));
      }
      umma_commit(smem_u32(&acc_full));
    }
    __syncwarp();
  } else {
    // ===================== element-wise warps: thread == (query row, 32-column quarter of the key tile) ===============
    const int lane_base = 32 * (warp & 3);
    const int row = lane_base + lane;
    const int quarter = (warp - 2) >> 2;
    const int cbase = quarter * 32;
    const uint32_t t_lane = (uint32_t)lane_base << 16;
    const float log2e = 1.4426950408889634f;
    float lse2 = 0.f, dsm = 0.f;
    if (kDQ) {
      const int qi = t0 + row;
      if (qi < p.L) {
        lse2 = p.lse[((size_t)b * p.heads + h) * p.L + qi] * log2e;
        dsm = p.dsum[((size_t)b * p.L + qi) * p.heads + h];
      }
    }
    uint8_t *srow_ds = smem_raw + (sdS - smem_u32(smem_raw)) + row * 16;
    uint8_t *srow_p = smem_raw + (sP - smem_u32(smem_raw)) + row * 16;
    // per-row scalars of the NEXT query tile are fetched one iteration ahead (their global-load latency would otherwise sit
    // between the accumulator wait and the arithmetic of every tile)
    float lse_n = 0.f, dsm_n = 0.f;
    auto fetch_row = [&](int qi, float &l2, float &dm) {
      l2 = 0.f; dm = 0.f;
      if (qi < p.L) {
        l2 = __ldg(p.lse + ((size_t)b * p.heads + h) * p.L + qi) * log2e;
        dm = __ldg(p.dsum + ((size_t)b * p.L + qi) * p.heads + h);
      }
    };
    if (!kDQ) fetch_row(row, lse_n, dsm_n);
    for (int it = 0; it < n_it; ++it) {
      if (!kDQ) {
        lse2 = lse_n; dsm = dsm_n;
        if (it + 1 < n_it) fetch_row((it + 1) * kBwT + row, lse_n, dsm_n);
      }
      const float dsm_s = dsm * p.scale;
      mbar_wait(smem_u32(&s_full), it & 1);
      tc_fence_after();
      uint32_t sr[32], dr[32];
#ifdef B200DIFF_DEBUG_TOGGLES
      if (p.debug & 4) {
#pragma unroll
        for (int i = 0; i < 32; ++i) { sr[i] = 0; dr[i] = 0; }
      } else
#endif
      {
        tmem_ld32(tmem_s + t_lane + cbase, sr);
        tmem_ld32(tmem_dp + t_lane + cbase, dr);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty));    // S / dP are in registers: the next tile's MMAs may overwrite them
      // keys past the end of the sequence (zero-filled by TMA) contribute nothing, but exp2(-lse) of their zero scores may
      // overflow the 16-bit range: mask them
      const int nvalid = min(kBwT, p.L - (kDQ ? it * kBwT : t0));
      uint4 pk_p[4], pk_s[4];
      if (nvalid == kBwT) {                              // every tile but (possibly) the last: no column masks
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          float pv[8], ds[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#ifdef B200DIFF_DEBUG_TOGGLES
            if (p.debug & 1) pv[i] = fmaf(__uint_as_float(sr[v8 * 8 + i]), p.scale_log2e, -lse2);
            else
#endif
            pv[i] = bw_ex2(fmaf(__uint_as_float(sr[v8 * 8 + i]), p.scale_log2e, -lse2));
            ds[i] = pv[i] * fmaf(__uint_as_float(dr[v8 * 8 + i]), p.scale, -dsm_s);
          }
          pk_s[v8] = pack8_fast<T>(ds);
          if (!kDQ) pk_p[v8] = pack8_fast<T>(pv);
        }
      } else {
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          float pv[8], ds[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pv[i] = bw_ex2(fmaf(__uint_as_float(sr[v8 * 8 + i]), p.scale_log2e, -lse2));
            if (cbase + v8 * 8 + i >= nvalid) pv[i] = 0.f;
            ds[i] = pv[i] * fmaf(__uint_as_float(dr[v8 * 8 + i]), p.scale, -dsm_s);
          }
          pk_s[v8] = pack8<T>(ds);
          if (!kDQ) pk_p[v8] = pack8<T>(pv);
        }
      }
      // the previous tile's dV / dK / dQ MMAs must have finished reading the operand tiles before they are rewritten
      mbar_wait(smem_u32(&p_empty), (it & 1) ^ 1);
#ifdef B200DIFF_DEBUG_TOGGLES
      if (!(p.debug & 2) || (pk_s[0].x == 0x12345678u && pk_p[0].x == 0x9abcdef0u))
#endif
#pragma unroll
      for (int v8 = 0; v8 < 4; ++v8) {
        const uint32_t off = ((cbase >> 3) + v8) * kBwBlk;
        *reinterpret_cast<uint4 *>(srow_ds + off) = pk_s[v8];
        if (!kDQ) *reinterpret_cast<uint4 *>(srow_p + off) = pk_p[v8];
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full));
    }
    // epilogue: accumulators -> dqkv (16-bit); the quarters split the D channels
    mbar_wait(smem_u32(&acc_full), 0);
    tc_fence_after();
    constexpr int kQD = D / 4;                           // 8 or 16 channels per quarter
    const int ti = t0 + row;
    T *orow = reinterpret_cast<T *>(p.dqkv) + ((size_t)b * p.L + ti) * p.ld + h * D + quarter * kQD;
#pragma unroll
    for (int c0 = 0; c0 < kQD; c0 += 8) {
      uint32_t r0[8], r1[8];
      tmem_ld8(tmem_a0 + t_lane + quarter * kQD + c0, r0);
      if (!kDQ) tmem_ld8(tmem_a1 + t_lane + quarter * kQD + c0, r1);
      tmem_ld_wait();
      if (ti < p.L) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r0[i]);
        *reinterpret_cast<uint4 *>(orow + (kDQ ? p.q_off : p.v_off) + c0) = pack8<T>(f);
        if (!kDQ) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r1[i]);
          *reinterpret_cast<uint4 *>(orow + p.k_off + c0) = pack8<T>(f);
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// v2 (head dim 64): TWO CTAs per SM.  The S / dP hand-off between the MMA issuer and the element-wise warps is a latency
// chain (commit -> mbarrier -> tcgen05.ld -> arrive -> next MMA) that one resident CTA cannot hide (tensor pipe 16 % active
// in v1, and switching off the exponentials, the operand stores and the second-stage MMAs one by one moved the time by < 15 %),
// so v2 halves every per-CTA resource: tiles of 128 x 64, TMEM 256 columns, ~96 KB of shared memory.
//   kDQ = true  : CTA = 128 queries (lanes), streams 64-KEY tiles:  S [128 q x 64 k] = Q K^T, dP = dO V^T, dQ += dS K.
//   kDQ = false : CTA = 128 keys (lanes), streams 64-QUERY tiles and works on the TRANSPOSED scores so that lanes stay keys:
//                 S^T [128 k x 64 q] = K Q^T, dP^T = V dO^T; P^T and dS^T are then K-major A operands (contraction over the
//                 queries) of dV += P^T dO and dK += dS^T Q; the per-query lse / D are per COLUMN (staged in shared memory).
// Every Q / K / V / dO tile is one SWIZZLE_128B TMA box; in both modes the second-stage operands are A = un-swizzled K-major
// tile written by the element-wise warps, B = the streamed tile read MN-major.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kV2Threads = 320;                          // warp 0 TMA, warp 1 MMA, warps 2-9 element-wise
constexpr int kV2S = 64;                                 // streamed tokens per tile

template <typename T, bool kDQ>
__global__ void __launch_bounds__(kV2Threads, 2)
attn_bwd_v2_kernel(const __grid_constant__ BwMaps maps, const BwParams p) {
  constexpr int D = 64;
  constexpr uint32_t kResTile = 128 * D * 2;             // 16 KB: 128 resident tokens
  constexpr uint32_t kStrTile = kV2S * D * 2;            // 8 KB: 64 streamed tokens
  constexpr uint32_t kOpTile = 128 * kV2S * 2;           // 16 KB: [128 lanes x 64] 16-bit operand tile (dS, or P^T / dS^T)
  constexpr int NS = kDQ ? 3 : 2;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t res_full, acc_full, s_full, s_empty, p_full, p_empty;
  __shared__ __align__(8) uint64_t str_full[NS], str_empty[NS];
  __shared__ uint32_t tmem_base_smem;
  // kDQ == false: per-column (query) scalars (lse * log2e, D * scale) of three consecutive tiles.  Tile it + 1's row is
  // written during iteration it BEFORE the writer's s_empty arrive; s_empty -> MMA issue -> s_full orders it before every
  // reader of iteration it + 1, and a warp can be at most one iteration ahead of another, so three rows never collide.
  __shared__ __align__(16) float2 s_col[3][kV2S];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_it = (p.L + kV2S - 1) / kV2S;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sRes0 = base, sRes1 = base + kResTile;
  const uint32_t sStr = sRes1 + kResTile;                // NS x (tile0, tile1)
  const uint32_t sOp0 = sStr + NS * 2 * kStrTile;        // dS (kDQ) / dS^T
  const uint32_t sOp1 = sOp0 + kOpTile;                  // P^T (only !kDQ)

  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.qkv); prefetch_tmap(&maps.dout); prefetch_tmap(&maps.qkv_s); prefetch_tmap(&maps.dout_s);
    mbar_init(smem_u32(&res_full), 1);
    mbar_init(smem_u32(&acc_full), 1);
    mbar_init(smem_u32(&s_full), 1);
    mbar_init(smem_u32(&s_empty), 8);
    mbar_init(smem_u32(&p_full), 8);
    mbar_init(smem_u32(&p_empty), 1);
    for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&str_full[i]), 1); mbar_init(smem_u32(&str_empty[i]), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 64, tmem_a0 = tmem_base + 128, tmem_a1 = tmem_base + 192;

  if (warp == 0) {
    if (lane == 0) {
      const int qc = p.q_off + h * D, kc = p.k_off + h * D, vc = p.v_off + h * D, oc = h * D;
      const uint32_t rb = smem_u32(&res_full);
      mbar_arrive_expect_tx(rb, 2 * kResTile);
      if (kDQ) {
        tma_load_3d(sRes0, &maps.qkv, rb, qc, t0, b);
        tma_load_3d(sRes1, &maps.dout, rb, oc, t0, b);
      } else {
        tma_load_3d(sRes0, &maps.qkv, rb, kc, t0, b);
        tma_load_3d(sRes1, &maps.qkv, rb, vc, t0, b);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it % NS;
        mbar_wait(smem_u32(&str_empty[st]), ((it / NS) & 1) ^ 1);
        const uint32_t bar = smem_u32(&str_full[st]);
        mbar_arrive_expect_tx(bar, 2 * kStrTile);
        const uint32_t d0 = sStr + st * 2 * kStrTile, d1 = d0 + kStrTile;
        if (kDQ) {
          tma_load_3d(d0, &maps.qkv_s, bar, kc, it * kV2S, b);
          tma_load_3d(d1, &maps.qkv_s, bar, vc, it * kV2S, b);
        } else {
          tma_load_3d(d0, &maps.qkv_s, bar, qc, it * kV2S, b);
          tma_load_3d(d1, &maps.dout_s, bar, oc, it * kV2S, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(smem_u32(&res_full), 0);
      // first stage: [128 resident lanes x 64 streamed columns], contraction over the 64 channels (+32 B per 16 channels)
      auto issue_first = [&](int g) {
        const int st = g % NS;
        mbar_wait(smem_u32(&str_full[st]), (g / NS) & 1);
        mbar_wait(smem_u32(&s_empty), (g & 1) ^ 1);
        tc_fence_after();
        const uint32_t x0 = sStr + st * 2 * kStrTile, x1 = x0 + kStrTile;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16(tmem_s, umma_desc_sw128(sRes0) + (uint64_t)(kk * 2), umma_desc_sw128(x0) + (uint64_t)(kk * 2), p.idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_f16(tmem_dp, umma_desc_sw128(sRes1) + (uint64_t)(kk * 2), umma_desc_sw128(x1) + (uint64_t)(kk * 2), p.idesc_s, kk > 0);
        umma_commit(smem_u32(&s_full));
      };
      issue_first(0);
      for (int it = 0; it < n_it; ++it) {
        if (it + 1 < n_it) issue_first(it + 1);
        const int st = it % NS;
        mbar_wait(smem_u32(&p_full), it & 1);
        tc_fence_after();
        const uint32_t x0 = sStr + st * 2 * kStrTile, x1 = x0 + kStrTile;
        const uint32_t acc = (it > 0) ? 1u : 0u;
        // second stage: contraction over the 64 streamed tokens.  A = operand tile [token block][lane][8 tokens] K-major
        // (16 tokens = 2 blocks = +4096 B); B = streamed tile read MN-major (8-token atoms 1024 B apart: +2048 B per 16)
#pragma unroll
        for (int kk = 0; kk < kV2S / 16; ++kk) {
          const uint64_t a0 = umma_desc_nosw(sOp0 + kk * 2 * kBwBlk, kBwBlk, 128);
          if (kDQ) {
            umma_f16(tmem_a0, a0, umma_desc_sw128(x0 + kk * 2048), p.idesc_acc, acc | (uint32_t)(kk > 0));        // dQ += dS K
          } else {
            umma_f16(tmem_a0, umma_desc_nosw(sOp1 + kk * 2 * kBwBlk, kBwBlk, 128), umma_desc_sw128(x1 + kk * 2048), p.idesc_acc,
                     acc | (uint32_t)(kk > 0));                                                                    // dV += P^T dO
            umma_f16(tmem_a1, a0, umma_desc_sw128(x0 + kk * 2048), p.idesc_acc, acc | (uint32_t)(kk > 0));        // dK += dS^T Q
          }
        }
        umma_commit(smem_u32(&str_empty[st]));
        umma_commit(smem_u32(&p_empty));
      }
      umma_commit(smem_u32(&acc_full));
    }
    __syncwarp();
  } else {
    // element-wise warps: thread == (lane row, 32-column half)
    const int lane_base = 32 * (warp & 3);
    const int row = lane_base + lane;
    const int half = (warp - 2) >> 2;
    const int cbase = half * 32;
    const int ew_tid = threadIdx.x - 64;                 // 0 .. 255
    const uint32_t t_lane = (uint32_t)lane_base << 16;
    const float log2e = 1.4426950408889634f;
    auto fetch = [&](int qi, float &l2, float &dm) {
      l2 = 0.f; dm = 0.f;
      if (qi < p.L) {
        l2 = __ldg(p.lse + ((size_t)b * p.heads + h) * p.L + qi) * log2e;
        dm = __ldg(p.dsum + ((size_t)b * p.L + qi) * p.heads + h) * p.scale;
      }
    };
    float lse2 = 0.f, dsm_s = 0.f;                       // kDQ: this thread's query row
    float nl = 0.f, nd = 0.f;                            // !kDQ: prefetched column scalars (threads 0..63 of the group)
    if (kDQ) fetch(t0 + row, lse2, dsm_s);
    else {
      if (ew_tid < kV2S) {
        fetch(ew_tid, nl, nd);
        s_col[0][ew_tid] = make_float2(nl, nd);
        if (n_it > 1) fetch(kV2S + ew_tid, nl, nd);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");      // once: row 0 is visible to all element-wise warps
    }
    uint8_t *srow0 = smem_raw + (sOp0 - smem_u32(smem_raw)) + row * 16;
    uint8_t *srow1 = smem_raw + (sOp1 - smem_u32(smem_raw)) + row * 16;
    for (int it = 0; it < n_it; ++it) {
      if (!kDQ && ew_tid < kV2S && it + 1 < n_it) {
        s_col[(it + 1) % 3][ew_tid] = make_float2(nl, nd);
        if (it + 2 < n_it) fetch((it + 2) * kV2S + ew_tid, nl, nd);
      }
      mbar_wait(smem_u32(&s_full), it & 1);
      tc_fence_after();
      uint32_t sr[32], dr[32];
      tmem_ld32(tmem_s + t_lane + cbase, sr);
      tmem_ld32(tmem_dp + t_lane + cbase, dr);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty));
      // dQ mode: streamed keys past the end of the sequence are masked (exp2(-lse) of their zero scores may overflow)
      const int nvalid = kDQ ? min(kV2S, p.L - it * kV2S) : kV2S;
      uint4 pk_p[4], pk_s[4];
#pragma unroll
      for (int v8 = 0; v8 < 4; ++v8) {
        float pv[8], ds[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const int c = cbase + v8 * 8 + i;
          float4 cc = make_float4(lse2, dsm_s, lse2, dsm_s);
          if (!kDQ) cc = *reinterpret_cast<const float4 *>(&s_col[it % 3][c]);      // two columns per 16-byte broadcast load
          pv[i] = bw_ex2(fmaf(__uint_as_float(sr[v8 * 8 + i]), p.scale_log2e, -cc.x));
          pv[i + 1] = bw_ex2(fmaf(__uint_as_float(sr[v8 * 8 + i + 1]), p.scale_log2e, -cc.z));
          if (kDQ && c >= nvalid) pv[i] = 0.f;
          if (kDQ && c + 1 >= nvalid) pv[i + 1] = 0.f;
          ds[i] = pv[i] * fmaf(__uint_as_float(dr[v8 * 8 + i]), p.scale, -cc.y);
          ds[i + 1] = pv[i + 1] * fmaf(__uint_as_float(dr[v8 * 8 + i + 1]), p.scale, -cc.w);
        }
        pk_s[v8] = pack8<T>(ds);
        if (!kDQ) pk_p[v8] = pack8<T>(pv);
      }
      mbar_wait(smem_u32(&p_empty), (it & 1) ^ 1);
#pragma unroll
      for (int v8 = 0; v8 < 4; ++v8) {
        const uint32_t off = ((cbase >> 3) + v8) * kBwBlk;
        *reinterpret_cast<uint4 *>(srow0 + off) = pk_s[v8];
        if (!kDQ) *reinterpret_cast<uint4 *>(srow1 + off) = pk_p[v8];
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full));
    }
    // epilogue: the halves split the 64 channels
    mbar_wait(smem_u32(&acc_full), 0);
    tc_fence_after();
    const int ti = t0 + row;
    T *orow = reinterpret_cast<T *>(p.dqkv) + ((size_t)b * p.L + ti) * p.ld + h * D + half * 32;
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 8) {
      uint32_t r0[8], r1[8];
      tmem_ld8(tmem_a0 + t_lane + half * 32 + c0, r0);
      if (!kDQ) tmem_ld8(tmem_a1 + t_lane + half * 32 + c0, r1);
      tmem_ld_wait();
      if (ti < p.L) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r0[i]);
        *reinterpret_cast<uint4 *>(orow + (kDQ ? p.q_off : p.v_off) + c0) = pack8<T>(f);
        if (!kDQ) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r1[i]);
          *reinterpret_cast<uint4 *>(orow + p.k_off + c0) = pack8<T>(f);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <typename T>
static int launch_bwd_v2(const void *qkv, int ld, int q_off, int k_off, int v_off, const void *dout, const float *lse,
                         const float *dsum, void *dqkv, int B, int L, int heads, float scale, cudaStream_t st) {
  constexpr int D = 64;
  BwMaps maps{};
  BwParams p{};
  const int ldo = heads * D;
  for (int which = 0; which < 4; ++which) {            // 128-token and 64-token boxes over qkv and dout
    const bool is_out = which & 1;
    const uint32_t rows = (which & 2) ? 64 : 128;
    const int pitch = is_out ? ldo : ld;
    uint64_t dims[3] = {(uint64_t)pitch, (uint64_t)L, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)pitch * 2, (uint64_t)L * pitch * 2};
    uint32_t box[3] = {64, rows, 1}, es[3] = {1, 1, 1};
    CUtensorMap *m = which == 0 ? &maps.qkv : which == 1 ? &maps.dout : which == 2 ? &maps.qkv_s : &maps.dout_s;
    if (int rc = cached_tmap(m, is_out ? dout : qkv, 3, dims, str, box, es, /*swizzle128=*/true)) return rc;
  }
  p.L = L; p.heads = heads; p.q_off = q_off; p.k_off = k_off; p.v_off = v_off; p.ld = ld;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.lse = lse; p.dsum = dsum; p.dqkv = dqkv; p.debug = 0;
  p.idesc_s = umma_idesc(Act<T>::kFmt, 128, 64, 0);
  p.idesc_acc = umma_idesc(Act<T>::kFmt, 128, 64, 1);       // A K-major, B MN-major
  const dim3 grid(cdiv(L, 128), heads, B);
  const size_t smem_kv = 2 * 16384 + 2 * 2 * 8192 + 2 * 16384 + 1024;
  const size_t smem_q = 2 * 16384 + 3 * 2 * 8192 + 16384 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_v2_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_v2_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_v2_kernel<T, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_v2_kernel<T, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  attn_bwd_v2_kernel<T, false><<<grid, kV2Threads, smem_kv, st>>>(maps, p);
  B200_LAUNCH_CHECK();
  attn_bwd_v2_kernel<T, true><<<grid, kV2Threads, smem_q, st>>>(maps, p);
  B200_LAUNCH_CHECK();
  return 0;
}

template <typename T, int D>
static int launch_bwd_tc(const void *qkv, int ld, int q_off, int k_off, int v_off, const void *dout, const float *lse,
                         const float *dsum, void *dqkv, int B, int L, int heads, float scale, cudaStream_t st) {
  BwMaps maps{};
  BwParams p{};
  {
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)L, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
    uint32_t box[3] = {(uint32_t)(D == 64 ? 64 : 8), 128, 1}, es[3] = {1, 1, 1};
    if (int rc = cached_tmap(&maps.qkv, qkv, 3, dims, str, box, es, /*swizzle128=*/D == 64)) return rc;
  }
  {
    const int ldo = heads * D;
    uint64_t dims[3] = {(uint64_t)ldo, (uint64_t)L, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)ldo * 2, (uint64_t)L * ldo * 2};
    uint32_t box[3] = {(uint32_t)(D == 64 ? 64 : 8), 128, 1}, es[3] = {1, 1, 1};
    if (int rc = cached_tmap(&maps.dout, dout, 3, dims, str, box, es, /*swizzle128=*/D == 64)) return rc;
  }
  p.L = L; p.heads = heads; p.q_off = q_off; p.k_off = k_off; p.v_off = v_off; p.ld = ld;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.lse = lse; p.dsum = dsum; p.dqkv = dqkv;
  p.debug = getenv("B200DIFF_ABW_DEBUG") ? atoi(getenv("B200DIFF_ABW_DEBUG")) : 0;
  p.idesc_s = umma_idesc(Act<T>::kFmt, 128, 128, 0);
  const size_t tile = (size_t)D * 256;
  const dim3 grid(cdiv(L, kBwT), heads, B);
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<T, D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<T, D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  // dK, dV: both second-stage operands MN-major (a_major bit 15, b_major bit 16)
  p.idesc_acc = umma_idesc(Act<T>::kFmt, 128, D, 1) | (1u << 15);
  attn_bwd_tc_kernel<T, D, false><<<grid, kBwThreads, tile * (2 + 2 * 3) + 2 * kBwPBytes + 1024, st>>>(maps, p);
  B200_LAUNCH_CHECK();
  // dQ: A = dS K-major, B = K MN-major
  p.idesc_acc = umma_idesc(Act<T>::kFmt, 128, D, 1);
  attn_bwd_tc_kernel<T, D, true><<<grid, kBwThreads, tile * (2 + 2 * 4) + kBwPBytes + 1024, st>>>(maps, p);
  B200_LAUNCH_CHECK();
  return 0;
}

int attention_bwd_tc(const void *qkv, int ld, int q_off, int k_off, int v_off, const void *dout, const float *lse,
                     const float *dsum, void *dqkv, int B, int L, int heads, int d, float scale, int act_dtype,
                     cudaStream_t st) {
  static const bool v1 = getenv("B200DIFF_ATTN_BWD_V1") != nullptr && getenv("B200DIFF_ATTN_BWD_V1")[0] == '1';
#define BW_(TT)                                                                                                         \
  do {                                                                                                                  \
    if (d == 32) return launch_bwd_tc<TT, 32>(qkv, ld, q_off, k_off, v_off, dout, lse, dsum, dqkv, B, L, heads, scale, st); \
    if (d == 64 && !v1) return launch_bwd_v2<TT>(qkv, ld, q_off, k_off, v_off, dout, lse, dsum, dqkv, B, L, heads, scale, st); \
    if (d == 64) return launch_bwd_tc<TT, 64>(qkv, ld, q_off, k_off, v_off, dout, lse, dsum, dqkv, B, L, heads, scale, st); \
  } while (0)
  if (act_dtype == B200_F16) BW_(__half);
  if (act_dtype == B200_BF16) BW_(__nv_bfloat16);
#undef BW_
  B200_CHECK(false, "attention_bwd: unsupported head dim %d / act_dtype %d", d, act_dtype);
}

}  // namespace b200
