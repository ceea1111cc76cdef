// Shared device/host helpers for libb200diff (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/b200diff.h"

namespace b200 {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define B200_CHECK(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      b200::set_error(__VA_ARGS__);      \
      return 1;                          \
    }                                    \
  } while (0)

#define B200_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      b200::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return 2;                                                                      \
    }                                                                                \
  } while (0)

// every launch goes through here so gpu_launches is a counted fact
#define B200_LAUNCH_CHECK()                                     \
  do {                                                          \
    b200::g_launches.fetch_add(1, std::memory_order_relaxed);   \
    B200_CUDA(cudaGetLastError());                              \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();

// ------------------------------------------------------------------------------------------
// 16-bit activation type traits
// ------------------------------------------------------------------------------------------
template <typename T>
struct Act;
template <>
struct Act<__half> {
  static constexpr int kFmt = 0;  // UMMA F16
  __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
  __device__ __forceinline__ static __half from_f(float v) {
    // saturate instead of producing inf: fp16 max is 65504
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    return __float2half_rn(v);
  }
};
template <>
struct Act<__nv_bfloat16> {
  static constexpr int kFmt = 1;  // UMMA BF16
  __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 8 x 16-bit packed in a uint4
template <typename T>
__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
  const T *p = reinterpret_cast<const T *>(&u);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = Act<T>::to_f(p[i]);
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  T *p = reinterpret_cast<T *>(&u);
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = Act<T>::from_f(f[i]);
  return u;
}

// x * sigmoid(x) with two MUFU ops (ex2.approx + rcp.approx, ~1e-7 relative) instead of an IEEE division
__device__ __forceinline__ float silu_f(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return v * r;
}
// SiLU of 8 values with 1.25 MUFU ops per element instead of 2: one ex2 each, and ONE reciprocal per four values —
// 1/a = (1/(abcd)) * bcd.  The GroupNorm+SiLU pass is HBM-bound on paper but sat at ~70 % MUFU-pipe utilisation (2 MUFU
// per element), which capped it at 0.72-0.82 of the copy bandwidth.  The exponent is clamped to 2^30 so the four-way
// product stays finite: sigmoid(v) for v < -20.8 becomes 9.3e-10 instead of <= 9.3e-10 (absolute error of SiLU < 2e-8).
// ~4 ulp (5e-7 relative) — three orders below the 16-bit rounding of the stored result.
__device__ __forceinline__ void silu8(float (&v)[8]) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v[i] * -1.4426950408889634f, 30.f)));
    a[i] = 1.f + e;
  }
#pragma unroll
  for (int h = 0; h < 8; h += 4) {
    const float pab = a[h] * a[h + 1], pcd = a[h + 2] * a[h + 3];
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(pab * pcd));
    const float rab = r * pcd, rcd = r * pab;
    v[h] *= rab * a[h + 1];
    v[h + 1] *= rab * a[h];
    v[h + 2] *= rcd * a[h + 3];
    v[h + 3] *= rcd * a[h + 2];
  }
}
// GELU(tanh) (DiTBlock's Mlp, UNet_DS_Diff/DiT_models.py:111): 0.5 x (1 + tanh(u)) = x * sigmoid(2u),
// u = sqrt(2/pi) (x + 0.044715 x^3); two MUFU ops (~1e-7 relative) instead of tanhf
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float u2 = 1.5957691216057308f * fmaf(0.044715f * x * x, x, x);     // 2u
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u2 * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}
// 8 fp32 -> 16-bit WITHOUT saturation (normalised / averaged values that cannot reach the fp16 range limit)
template <typename T>
__device__ __forceinline__ uint4 pack8_fast(const float (&f)[8]);
template <>
__device__ __forceinline__ uint4 pack8_fast<__half>(const float (&f)[8]) {
  uint4 u;
  __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
template <>
__device__ __forceinline__ uint4 pack8_fast<__nv_bfloat16>(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// two fp32 -> one packed 16-bit pair with saturation to the largest finite value (single F2FP.SATFINITE)
template <typename T>
__device__ __forceinline__ uint32_t pack2_sat(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2_sat<__half>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack2_sat<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <typename T>
__device__ __forceinline__ uint4 pack8_sat(const float *f) {
  return make_uint4(pack2_sat<T>(f[0], f[1]), pack2_sat<T>(f[2], f[3]), pack2_sat<T>(f[4], f[5]), pack2_sat<T>(f[6], f[7]));
}
// one packed 16-bit pair -> two fp32
template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t w);
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// streaming 128-bit access (read-once / write-once tensors)
__device__ __forceinline__ uint4 ld_stream(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void *p, const uint4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if ((++spins & 0x3ff) == 0) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {  // ~2 s at 2 GHz
        printf("b200diff: mbarrier wait timeout (block %d thread %d bar %u parity %u)\n",
               blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void prefetch_tmap(const void *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void *tmap, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void *tmap, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address to smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets row (lane_base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Per-(image, channel) statistics epilogue of the HBM-bound producers: the `rows` threads of a block that own the same
// 8-channel column are reduced through shared memory, then ONE fp64 atomic pair per channel per block goes to the global
// sums (thread layout: threadIdx.x = row * ncol + col, 256 threads).  Must be called by every thread of the block.
__device__ __forceinline__ void block_flush_col_sums(const float (&s)[8], const float (&q)[8], bool active, int row, int col,
                                                     int ncol, int rows, double *dst /* sums + (n*C + col*8)*2 */) {
  __shared__ float red[256 * 16];
  if (active) {
    float *r = red + (row * ncol + col) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[i] = s[i]; r[8 + i] = q[i]; }
  }
  __syncthreads();
  if (active && row == 0) {
    float ts[8], tq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ts[i] = s[i]; tq[i] = q[i]; }
    for (int rr = 1; rr < rows; ++rr) {
      const float *r = red + (rr * ncol + col) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) { ts[i] += r[i]; tq[i] += r[8 + i]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(dst + 2 * i, (double)ts[i]);
      atomicAdd(dst + 2 * i + 1, (double)tq[i]);
    }
  }
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per thread
__device__ __forceinline__ void ld_global_256(const void *p, uint4 &a, uint4 &b) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st_global_256(void *p, const uint4 &a, const uint4 &b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// wait for the outstanding tcgen05.ld and tie the destination registers to the wait, so no use of them can be
// scheduled above it when loads are issued ahead of their consumer
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// registers -> TMEM: 32 lanes x 8 consecutive fp32 columns (accumulator rescale of the online softmax)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Un-swizzled ("interleaved") operand tile made of 8-row x 16-byte core matrices (128 contiguous bytes each).
// lbo = byte distance between core matrices adjacent along K, sbo = along M/N.  Valid for K-major operands
// ([row][8 elems] core rows) and MN-major ones ([k][8 mn-elems] core rows) alike.
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo & 0x3FFFF) >> 4) << 16;
  d |= (uint64_t)((sbo & 0x3FFFF) >> 4) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  return d;                               // layout_type 0 = SWIZZLE_NONE
}

// K-major, 128B-swizzled operand tile: rows of 64 x 16-bit (=128 B), 8-row atoms 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
//  | layout_type [61,64) with SWIZZLE_128B = 2.)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)0 << 16;                 // LBO unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16, fp32 accumulate, both operands K-major.
__host__ __device__ inline uint32_t umma_idesc(int fmt /*0 f16, 1 bf16*/, int M, int N, int b_mn_major = 0) {
  uint32_t d = 0;
  d |= (uint32_t)(b_mn_major & 1) << 16;  // b_major: 0 = K-major, 1 = MN-major
  d |= 1u << 4;                       // c_format = F32
  d |= (uint32_t)fmt << 7;            // a_format
  d |= (uint32_t)fmt << 10;           // b_format
  d |= (uint32_t)(N >> 3) << 17;      // n_dim
  d |= (uint32_t)(M >> 4) << 24;      // m_dim
  return d;
}

// ------------------------------------------------------------------------------------------
// host: TMA descriptor encode through the runtime-resolved driver entry point
// ------------------------------------------------------------------------------------------
int encode_tmap(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                const uint64_t *strides_bytes /*rank-1*/, const uint32_t *box,
                const uint32_t *elem_strides, bool swizzle128 = true);
int cached_tmap(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                const uint64_t *strides_bytes, const uint32_t *box, const uint32_t *elem_strides,
                bool swizzle128 = true);

}  // namespace b200
