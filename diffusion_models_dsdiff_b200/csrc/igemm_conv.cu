// Implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::f16, fp32 accumulate in TMEM), operands
// staged by TMA (tiled 4-D boxes over NHWC activations, OOB zero-fill = conv padding; strided
// boxes = stride-2 conv), persistent warp-specialised CTAs with a double-buffered accumulator.
//
//   GEMM view:  D[pixel, cout] = sum_{seg, tap, cin} A_seg[pixel shifted by tap, cin] * W_seg[tap][cout][cin]
//   M = 128 output pixels (a TW x TH x TN box), N = BN output channels (<= 256), K chunk = 64 channels.
//
// Replaces F.conv2d at the call sites listed in include/b200diff.h (b200_conv2d_igemm).
#include "common.cuh"

namespace b200 {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kMaxStages = 8;
constexpr int kConvThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kAccCols = 256;      // TMEM columns per accumulator stage
constexpr int kTmemCols = 512;     // two accumulator stages
constexpr int kABytes = kBM * kBK * 2;

struct ConvMaps {
  CUtensorMap a[3];
  CUtensorMap b[3];
};

struct ConvKParams {
  int N, Hout, Wout, Cout;
  int BN, n_ntiles;
  int TW, TH, TN;
  int tiles_w, tiles_h, tiles_n;
  int num_tiles;
  int stride;
  int stages;
  int nseg;
  int seg_cin[3], seg_coff[3], seg_ks[3], seg_pad[3];
  const float *bias;
  const float *rowbias;
  int rowbias_stride;
  const void *residual;
  float res_scale;
  void *out;
  double *gn_sums;
  uint32_t idesc;
};

// Column sums of a 32 (lanes) x 16 (registers) tile by recursive halving: 16 shuffles instead of 16 x 5.
// Returns on every lane the total of column (lane >> 1).
__device__ __forceinline__ float warp_colsum16(const float (&a)[16], int lane) {
  float b[8];
  const bool h16 = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h16 ? a[i + 8] : a[i];
    const float send = h16 ? a[i] : a[i + 8];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float c[4];
  const bool h8 = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h8 ? b[i + 4] : b[i];
    const float send = h8 ? b[i] : b[i + 4];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float d[2];
  const bool h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h4 ? c[i + 2] : c[i];
    const float send = h4 ? c[i] : c[i + 2];
    d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const bool h2 = lane & 2;
  const float keep = h2 ? d[1] : d[0];
  const float send = h2 ? d[0] : d[1];
  float e = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  e += __shfl_xor_sync(0xffffffffu, e, 1);
  return e;
}

template <typename T>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvMaps maps, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + (uint32_t)p.stages * kABytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      prefetch_tmap(&maps.a[s]);
      prefetch_tmap(&maps.b[s]);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_ntiles;
        int pt = tile / p.n_ntiles;
        const int w0 = (pt % p.tiles_w) * p.TW;
        pt /= p.tiles_w;
        const int h0 = (pt % p.tiles_h) * p.TH;
        const int n0 = (pt / p.tiles_h) * p.TN;
        for (int s = 0; s < p.nseg; ++s) {
          const int ks = p.seg_ks[s], pad = p.seg_pad[s], cin = p.seg_cin[s], coff = p.seg_coff[s];
          for (int tap = 0; tap < ks * ks; ++tap) {
            const int r = tap / ks, sx = tap - r * ks;
            const int hin = h0 * p.stride + r - pad;
            const int win = w0 * p.stride + sx - pad;
            for (int c0 = 0; c0 < cin; c0 += kBK) {
              mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
              const uint32_t fb = smem_u32(&full_bar[stage]);
              mbar_arrive_expect_tx(fb, kABytes + b_bytes);
              tma_load_4d(smem_a + stage * kABytes, &maps.a[s], fb, c0, win, hin, n0);
              tma_load_3d(smem_b + stage * b_bytes, &maps.b[s], fb, coff + c0, n_tile * p.BN, tap);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccCols;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_ks[s] * p.seg_ks[s], cin = p.seg_cin[s];
          for (int tap = 0; tap < taps; ++tap) {
            for (int c0 = 0; c0 < cin; c0 += kBK) {
              const int kk = min(kBK, cin - c0);
              const int nk = (kk + 15) >> 4;
              mbar_wait(smem_u32(&full_bar[stage]), phase);
              tc_fence_after();
              const uint64_t da = umma_desc_sw128(smem_a + stage * kABytes);
              const uint64_t db = umma_desc_sw128(smem_b + stage * b_bytes);
              for (int k = 0; k < nk; ++k) {
                // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in >>4 units
                umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, accumulate);
                accumulate = 1;
              }
              umma_commit(smem_u32(&empty_bar[stage]));
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
        umma_commit(smem_u32(&tmem_full_bar[acc]));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (4 warps = 128 TMEM lanes) =====================
    const int lane_base = 32 * (warp & 3);
    const int m = lane_base + lane;
    const int tw = m % p.TW;
    const int th = (m / p.TW) % p.TH;
    const int tn = m / (p.TW * p.TH);
    T *out = reinterpret_cast<T *>(p.out);
    const T *res = reinterpret_cast<const T *>(p.residual);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile = tile % p.n_ntiles;
      int pt = tile / p.n_ntiles;
      const int w = (pt % p.tiles_w) * p.TW + tw;
      pt /= p.tiles_w;
      const int h = (pt % p.tiles_h) * p.TH + th;
      const int n0 = (pt / p.tiles_h) * p.TN;
      const int n = n0 + tn;
      const int n_warp = n0 + lane_base / (p.TW * p.TH);
      const bool valid = (w < p.Wout) && (h < p.Hout) && (n < p.N);
      const size_t pix = ((size_t)n * p.Hout + h) * p.Wout + w;
      mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + acc * kAccCols;
      for (int c = 0; c < p.BN; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c, r);
        tmem_ld_wait();
        const int co = n_tile * p.BN + c;
        float v[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int cb = co + half * 8;
          float (&vh)[8] = *reinterpret_cast<float (*)[8]>(v + half * 8);
#pragma unroll
          for (int i = 0; i < 8; ++i) vh[i] = __uint_as_float(r[half * 8 + i]);
          const bool live = valid && cb < p.Cout;
          if (live) {
            if (p.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4 *>(p.bias + cb));
              const float4 b1 = __ldg(reinterpret_cast<const float4 *>(p.bias + cb + 4));
              vh[0] += b0.x; vh[1] += b0.y; vh[2] += b0.z; vh[3] += b0.w;
              vh[4] += b1.x; vh[5] += b1.y; vh[6] += b1.z; vh[7] += b1.w;
            }
            if (p.rowbias) {
              const float *rb = p.rowbias + (size_t)n * p.rowbias_stride + cb;
              const float4 b0 = __ldg(reinterpret_cast<const float4 *>(rb));
              const float4 b1 = __ldg(reinterpret_cast<const float4 *>(rb + 4));
              vh[0] += b0.x; vh[1] += b0.y; vh[2] += b0.z; vh[3] += b0.w;
              vh[4] += b1.x; vh[5] += b1.y; vh[6] += b1.z; vh[7] += b1.w;
            }
            if (res) {
              const uint4 ru = *reinterpret_cast<const uint4 *>(res + pix * p.Cout + cb);
              float rf[8];
              unpack8<T>(ru, rf);
#pragma unroll
              for (int i = 0; i < 8; ++i) vh[i] += p.res_scale * rf[i];
            }
            *reinterpret_cast<uint4 *>(out + pix * p.Cout + cb) = pack8<T>(vh);
          } else if (p.gn_sums) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vh[i] = 0.f;
          }
        }
        if (p.gn_sums) {
          // per-channel (sum, sum of squares) over this warp's 32 pixels for a following GroupNorm; the 32 rows of a
          // warp lie in ONE image (the host enables the fused statistics only when TW*TH >= 32)
          float q[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) q[i] = v[i] * v[i];
          const float s_tot = warp_colsum16(v, lane);
          const float q_tot = warp_colsum16(q, lane);
          const int ch = co + (lane >> 1);
          if ((lane & 1) == 0 && ch < p.Cout && n_warp < p.N) {
            double *dst = p.gn_sums + ((size_t)n_warp * p.Cout + ch) * 2;
            atomicAdd(dst, (double)s_tot);
            atomicAdd(dst + 1, (double)q_tot);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace {
int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

// largest N tile <= 256 that divides Cout (multiple of 16); falls back to padded tiles
int pick_bn(int Cout) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (Cout % bn == 0) return bn;
  return Cout < 256 ? ((Cout + 15) / 16) * 16 : 256;
}
}  // namespace

template <typename T>
static int launch_conv(const b200_conv_desc *d, cudaStream_t stream) {
  ConvKParams p{};
  ConvMaps maps{};
  p.N = d->N; p.Hout = d->Hout; p.Wout = d->Wout; p.Cout = d->Cout;
  p.stride = d->stride;
  p.nseg = d->nseg;
  p.BN = pick_bn(d->Cout);
  p.n_ntiles = cdiv(d->Cout, p.BN);
  // pixel box: TW x TH x TN == 128.  The box is chosen from (Hout, Wout) ONLY — never from N — so that the result
  // for one image is bit-identical whatever batch it is part of (slices are independent units):
  // minimise the padded area of one image, prefer 16-wide boxes on ties (square-ish halo => better L2 locality).
  long best = -1;
  for (int tw = 32; tw >= 1; tw >>= 1) {
    int th = std::min(pow2_ceil(d->Hout), kBM / tw);
    int tn = kBM / (tw * th);
    long area = (long)cdiv(d->Wout, tw) * tw * cdiv(d->Hout, th) * th;
    if (best < 0 || area < best || (area == best && tw == 16)) {
      best = area; p.TW = tw; p.TH = th; p.TN = tn;
    }
  }
  p.tiles_w = cdiv(d->Wout, p.TW);
  p.tiles_h = cdiv(d->Hout, p.TH);
  p.tiles_n = cdiv(d->N, p.TN);
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_ntiles;
  p.bias = d->bias; p.rowbias = d->rowbias; p.rowbias_stride = d->rowbias_stride;
  p.residual = d->residual; p.res_scale = d->res_scale; p.out = d->out;
  // fused statistics need every epilogue warp (32 consecutive box rows) inside one image; tinier maps (< 32 pixels)
  // take the standalone kernel below — a choice that depends on the map size only, not on N
  p.gn_sums = (p.TW * p.TH >= 32) ? d->gn_sums : nullptr;
  p.idesc = umma_idesc(Act<T>::kFmt, kBM, p.BN);
  const size_t stage_bytes = (size_t)kABytes + (size_t)p.BN * 128;
  p.stages = (int)std::min<size_t>(kMaxStages, (200 * 1024) / stage_bytes);
  B200_CHECK(p.stages >= 2, "conv: tile too large for shared memory");

  for (int s = 0; s < d->nseg; ++s) {
    const b200_conv_seg &g = d->seg[s];
    B200_CHECK(g.ksize == 1 || g.ksize == 3, "conv: ksize must be 1 or 3 (got %d)", g.ksize);
    B200_CHECK(g.cin % 8 == 0 && g.wcin % 8 == 0, "conv: channel counts must be multiples of 8");
    B200_CHECK(g.coff % 8 == 0 && g.coff + g.cin <= g.wcin, "conv: bad weight channel window");
    B200_CHECK(((uintptr_t)g.src & 15) == 0 && ((uintptr_t)g.weight & 15) == 0, "conv: unaligned pointer");
    p.seg_cin[s] = g.cin; p.seg_coff[s] = g.coff; p.seg_ks[s] = g.ksize; p.seg_pad[s] = g.pad;
    {  // activations: (C, W, H, N)
      uint64_t dims[4] = {(uint64_t)g.cin, (uint64_t)d->Win, (uint64_t)d->Hin, (uint64_t)d->N};
      uint64_t str[3] = {(uint64_t)g.cin * 2, (uint64_t)d->Win * g.cin * 2,
                         (uint64_t)d->Hin * d->Win * g.cin * 2};
      uint32_t box[4] = {kBK, (uint32_t)(p.TW * d->stride), (uint32_t)(p.TH * d->stride), (uint32_t)p.TN};
      uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
      if (int rc = cached_tmap(&maps.a[s], g.src, 4, dims, str, box, es)) return rc;
    }
    {  // weights: (Cin, Cout, taps)
      uint64_t dims[3] = {(uint64_t)g.wcin, (uint64_t)d->Cout, (uint64_t)(g.ksize * g.ksize)};
      uint64_t str[2] = {(uint64_t)g.wcin * 2, (uint64_t)d->Cout * g.wcin * 2};
      uint32_t box[3] = {kBK, (uint32_t)p.BN, 1};
      uint32_t es[3] = {1, 1, 1};
      if (int rc = cached_tmap(&maps.b[s], g.weight, 3, dims, str, box, es)) return rc;
    }
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  static bool attr_set[2] = {false, false};
  if (!attr_set[Act<T>::kFmt]) {
    B200_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   220 * 1024));
    attr_set[Act<T>::kFmt] = true;
  }
  const int grid = std::min(p.num_tiles, num_sms());
  conv_igemm_kernel<T><<<grid, kConvThreads, smem, stream>>>(maps, p);
  B200_LAUNCH_CHECK();
  if (d->gn_sums && !p.gn_sums)
    return b200_channel_stats(d->out, d->N, d->Hout * d->Wout, d->Cout, d->gn_sums, d->act_dtype, stream);
  return 0;
}

}  // namespace b200

extern "C" int b200_conv2d_igemm(const b200_conv_desc *d, void *stream) {
  using namespace b200;
  B200_CHECK(d != nullptr, "conv: null descriptor");
  B200_CHECK(d->nseg >= 1 && d->nseg <= 3, "conv: nseg must be 1..3");
  B200_CHECK(d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  B200_CHECK(d->Cout % 8 == 0, "conv: Cout must be a multiple of 8 (got %d)", d->Cout);
  B200_CHECK(d->N > 0 && d->Hout > 0 && d->Wout > 0, "conv: empty output");
  if (d->act_dtype == B200_F16) return launch_conv<__half>(d, (cudaStream_t)stream);
  if (d->act_dtype == B200_BF16) return launch_conv<__nv_bfloat16>(d, (cudaStream_t)stream);
  B200_CHECK(false, "conv: bad act_dtype %d", d->act_dtype);
}
