// Weight-gradient GEMM of a Linear layer straight from the token-major tensors (training step):
//     dW[O, I] = sum_t dY[t, O] * X[t, I]          (what loss.backward() computes for nn.Linear.weight;
//                                                    UNet_DS_Diff/DiT_models.py:101-142 via timm Attention / Mlp)
// Both operands have their GEMM K dimension (tokens) as the OUTER, strided axis, so for tcgen05 they are MN-major:
// a TMA box of (64 channels x 64 tokens) with SWIZZLE_128B lands as the canonical MN-major SW128 tile (row = one token =
// 128 bytes of 64 channels; 8-token atoms 1024 B apart along K; the next 64 channels are the next box, 8192 B along M / N),
// and the instruction descriptor's a_major / b_major bits tell the tensor core to contract over the rows.  No transposed
// copy of dY or X is ever written (the first version transposed both with an extra HBM pass each and ran the K-major
// implicit-GEMM kernel on the copies).
//
// Persistent CTAs walk (128 x 256 output tile, K split) work items: warp 0 = TMA producer (4-stage ring of 16 KB A + 32 KB
// B), warp 1 = MMA issuer (4 x M128 N256 K16 per stage) with two 256-column TMEM accumulators so that the epilogue of item i
// overlaps the MMAs of item i + 1, warps 2-5 = epilogue (tcgen05.ld -> fp32 stores).  K splits write fp32 partial planes
// that a second kernel adds in fixed order (deterministic, like the implicit-GEMM kernel's split-K).
#include "common.cuh"

namespace b200 {

constexpr int kWgThreads = 192;
constexpr int kWgBM = 128, kWgBN = 256, kWgBK = 64;
constexpr int kWgStages = 4;
constexpr uint32_t kWgBox = 64 * 128;                    // (64 channels x 64 tokens) x 2 B = 8192
constexpr uint32_t kWgABytes = 2 * kWgBox, kWgBBytes = 4 * kWgBox;
constexpr uint32_t kWgStageBytes = kWgABytes + kWgBBytes;

struct WgMaps {
  CUtensorMap a, b;
};

struct WgParams {
  int M, N, T;                  // dW is [M, N]; T tokens
  int tiles_m, tiles_n, splits;
  int chunks;                   // K chunks of 64 tokens in total
  float *out;                   // [splits][M][N] fp32 (splits == 1: the result itself)
  uint32_t idesc;
};

// MN-major SWIZZLE_128B operand: 8-row atoms 1024 B apart along K (SBO); `lbo` = byte distance between the 64-element atoms
// along M / N
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo & 0x3FFFF) >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <typename T>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tn_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgStages], empty_bar[kWgStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int items = p.tiles_m * p.tiles_n * p.splits;

  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.a); prefetch_tmap(&maps.b);
    for (int i = 0; i < kWgStages; ++i) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  auto decode = [&](int item, int &m0, int &n0, int &c_begin, int &c_end) {
    const int split = item % p.splits;
    const int tile = item / p.splits;
    m0 = (tile % p.tiles_m) * kWgBM;
    n0 = (tile / p.tiles_m) * kWgBN;
    c_begin = (int)((long long)split * p.chunks / p.splits);
    c_end = (int)((long long)(split + 1) * p.chunks / p.splits);
  };

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int m0, n0, cb, ce;
        decode(item, m0, n0, cb, ce);
        for (int c = cb; c < ce; ++c) {
          mbar_wait(smem_u32(&empty_bar[st]), ph ^ 1);
          const uint32_t bar = smem_u32(&full_bar[st]);
          const uint32_t sa = base + st * kWgStageBytes, sb = sa + kWgABytes;
          mbar_arrive_expect_tx(bar, kWgStageBytes);
          for (int j = 0; j < 2; ++j) tma_load_3d(sa + j * kWgBox, &maps.a, bar, m0 + 64 * j, c * kWgBK, 0);
          for (int j = 0; j < 4; ++j) tma_load_3d(sb + j * kWgBox, &maps.b, bar, n0 + 64 * j, c * kWgBK, 0);
          if (++st == kWgStages) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int st = 0, it = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        int m0, n0, cb, ce;
        decode(item, m0, n0, cb, ce);
        const int acc = it & 1;
        mbar_wait(smem_u32(&acc_empty[acc]), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int c = cb; c < ce; ++c) {
          mbar_wait(smem_u32(&full_bar[st]), ph);
          tc_fence_after();
          const uint32_t sa = base + st * kWgStageBytes, sb = sa + kWgABytes;
#pragma unroll
          for (int kk = 0; kk < kWgBK / 16; ++kk)       // 16 tokens = two 8-row atoms = +2048 B
            umma_f16(tmem_d, umma_desc_sw128_mn(sa + kk * 2048, kWgBox), umma_desc_sw128_mn(sb + kk * 2048, kWgBox), p.idesc,
                     (c > cb || kk > 0) ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[st]));
          if (++st == kWgStages) { st = 0; ph ^= 1; }
        }
        umma_commit(smem_u32(&acc_full[acc]));
      }
    }
    __syncwarp();
  } else {
    // epilogue warps 2-5: TMEM lane quadrant = warp % 4
    const int lane_base = 32 * (warp & 3);
    const int row = lane_base + lane;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      int m0, n0, cb, ce;
      decode(item, m0, n0, cb, ce);
      const int acc = it & 1;
      mbar_wait(smem_u32(&acc_full[acc]), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + acc * 256;
      float *dst = p.out + (size_t)(item % p.splits) * p.M * p.N + (size_t)(m0 + row) * p.N + n0;
      const bool row_ok = m0 + row < p.M;
#pragma unroll 4
      for (int c = 0; c < kWgBN; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c, r);
        tmem_ld_wait();
        if (row_ok && n0 + c < p.N) {
          float4 *d4 = reinterpret_cast<float4 *>(dst + c);
          d4[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
          d4[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
          d4[2] = make_float4(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
          d4[3] = make_float4(__uint_as_float(r[12]), __uint_as_float(r[13]), __uint_as_float(r[14]), __uint_as_float(r[15]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[acc]));
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ ws, int S, long long plane, float *__restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = *reinterpret_cast<const float4 *>(ws + i * 4);
    for (int k = 1; k < S; ++k) {
      const float4 b = *reinterpret_cast<const float4 *>(ws + (size_t)k * plane + i * 4);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4 *>(out + i * 4) = a;
  }
}

template <typename T>
static int launch_wgrad(const void *dy, const void *x, float *out, float *ws, int64_t Tn, int M, int N, int splits,
                        cudaStream_t st) {
  WgMaps maps{};
  WgParams p{};
  {
    uint64_t dims[3] = {(uint64_t)M, (uint64_t)Tn, 1};
    uint64_t str[2] = {(uint64_t)M * 2, (uint64_t)Tn * M * 2};
    uint32_t box[3] = {64, 64, 1}, es[3] = {1, 1, 1};
    if (int rc = cached_tmap(&maps.a, dy, 3, dims, str, box, es, /*swizzle128=*/true)) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)Tn, 1};
    uint64_t str[2] = {(uint64_t)N * 2, (uint64_t)Tn * N * 2};
    uint32_t box[3] = {64, 64, 1}, es[3] = {1, 1, 1};
    if (int rc = cached_tmap(&maps.b, x, 3, dims, str, box, es, /*swizzle128=*/true)) return rc;
  }
  p.M = M; p.N = N; p.T = (int)Tn;
  p.tiles_m = cdiv(M, kWgBM); p.tiles_n = cdiv(N, kWgBN); p.splits = splits;
  p.chunks = (int)cdiv64(Tn, kWgBK);
  p.out = splits > 1 ? ws : out;
  p.idesc = umma_idesc(Act<T>::kFmt, kWgBM, kWgBN, 1) | (1u << 15);      // a_major = b_major = MN
  const size_t smem = (size_t)kWgStages * kWgStageBytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(wgrad_tn_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int items = p.tiles_m * p.tiles_n * splits;
  wgrad_tn_kernel<T><<<std::min(items, num_sms()), kWgThreads, smem, st>>>(maps, p);
  B200_LAUNCH_CHECK();
  if (splits > 1) {
    const long long plane = (long long)M * N, n4 = plane / 4;
    wgrad_reduce_kernel<<<(int)std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 8), 256, 0, st>>>(ws, splits, plane,
                                                                                                              out, n4);
    B200_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_wgrad_splits(int64_t T, int M, int N) {
  // (tile, split) items for the persistent CTAs: the smallest split whose last wave is at least 85 % full, every split at
  // least 8 chunks of 64 tokens long.  Geometry only.
  if (T <= 0 || M <= 0 || N <= 0) return 1;
  const long long tiles = (long long)cdiv(M, kWgBM) * cdiv(N, kWgBN);
  const int chunks = (int)cdiv64(T, kWgBK);
  const int sms = num_sms();
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 16 && chunks / s >= 8; ++s) {
    const long long items = tiles * s;
    const long long waves = (items + sms - 1) / sms;
    const double eff = (double)items / (double)(waves * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    if (eff >= 0.85) { best = s; break; }
  }
  return best;
}

extern "C" int b200_linear_wgrad(const void *dy, const void *x, float *out, float *workspace, int64_t T, int M, int N,
                                 int act_dtype, void *stream) {
  B200_CHECK(dy && x && out && T > 0 && M > 0 && N > 0, "linear_wgrad: bad arguments");
  B200_CHECK(M % 64 == 0 && N % 64 == 0 && T % 8 == 0, "linear_wgrad: M and N must be multiples of 64, T of 8 (got %d, %d, %lld)", M,
             N, (long long)T);
  B200_CHECK((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)out) & 15) == 0, "linear_wgrad: unaligned pointer");
  const int splits = b200_wgrad_splits(T, M, N);
  B200_CHECK(splits == 1 || workspace != nullptr, "linear_wgrad: %d K splits need a workspace of splits * M * N floats", splits);
  if (act_dtype == B200_F16) return launch_wgrad<__half>(dy, x, out, workspace, T, M, N, splits, (cudaStream_t)stream);
  if (act_dtype == B200_BF16) return launch_wgrad<__nv_bfloat16>(dy, x, out, workspace, T, M, N, splits, (cudaStream_t)stream);
  B200_CHECK(false, "linear_wgrad: bad act_dtype");
}
