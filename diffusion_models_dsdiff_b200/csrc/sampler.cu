// Fused per-step sampler update: eps/v -> x0 prediction, clip, posterior mean / DDIM direction,
// variance and noise add in ONE elementwise kernel (HBM-bound, float4 vectorised).  The step's
// coefficient row is staged into shared memory by every CTA from a device table indexed by a
// device-resident step counter, so one captured CUDA graph serves every step.
//
// Arithmetic follows the reference's operation order with explicit round-to-nearest mul/add (no FMA
// contraction) so fp32 results agree with the PyTorch path to the last bit or two:
//   LDM_DDIM  ldm/models/diffusion/ddim.py:221-261 (+ ddpm.py:290-302 for v)
//   LDM_DDPM  trainers/trainer_ds_diff.py:717-738, ldm/models/diffusion/ddpm.py:304-332
//   GD_DDIM   training_project/utils/gaussian_diffusion.py:620-667 (p_mean_variance :247-351)
//   GD_DDPM   training_project/utils/gaussian_diffusion.py:423-466
//   DPMPP_2M  ldm/models/diffusion/dpm_solver_new/dpm_solver_pytorch.py:284-300, 435-444, 549-575, 798-835
#include "common.cuh"

namespace b200 {

struct SamplerK {
  int mode, param, clip, learned_range;
  int C, HW;           // per-sample channel count and plane size
  long long total;     // N*C*HW
  const float *x, *m, *noise, *coef;
  const int *step_ptr;
  int step_host;
  float *x_prev, *pred_x0;
  float *hist;
  const float *thresh;
  int phase;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ void sampler_elem(const SamplerK &k, const float *c, float x, float m,
                                             float s, float z, float &xp, float &x0o) {
  float x0, out;
  if (k.mode == B200_SAMPLER_LDM_DDIM) {
    float e;
    if (k.param == 2) {  // v
      e = add(mul(c[0], m), mul(c[1], x));
      x0 = sub(mul(c[0], x), mul(c[1], m));
    } else {
      e = m;
      x0 = __fdiv_rn(sub(x, mul(c[3], e)), c[2]);
    }
    if (k.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    const float dir = mul(c[5], e);
    const float nz = mul(mul(c[6], z), c[7]);
    out = add(add(mul(c[4], x0), dir), nz);
  } else if (k.mode == B200_SAMPLER_LDM_DDPM) {
    x0 = (k.param == 1) ? m : sub(mul(c[0], x), mul(c[1], m));
    if (k.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    const float mean = add(mul(c[2], x0), mul(c[3], x));
    out = add(mean, mul(c[4], z));
  } else {
    x0 = (k.param == 1) ? m : sub(mul(c[0], x), mul(c[1], m));
    if (k.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    if (k.mode == B200_SAMPLER_GD_DDIM) {
      const float eps = __fdiv_rn(sub(mul(c[6], x), x0), c[7]);
      const float mean_pred = add(mul(x0, c[8]), mul(c[9], eps));
      out = add(mean_pred, mul(c[10], z));
    } else {
      const float mean = add(mul(c[2], x0), mul(c[3], x));
      float logvar;
      if (k.learned_range) {
        const float frac = __fdiv_rn(add(s, 1.f), 2.f);
        logvar = add(mul(frac, c[5]), mul(sub(1.f, frac), c[4]));
      } else {
        logvar = c[4];
      }
      out = add(mean, mul(mul(c[11], expf(mul(0.5f, logvar))), z));
    }
  }
  xp = out;
  x0o = x0;
}

// DPM-Solver++ 2M.  c: 0 alpha_s, 1 sigma_s, 2 sigma_t/sigma_s, 3 alpha_t*phi_1, 4 1/r0, 5 0.5*alpha_t*phi_1, 6 order.
__device__ __forceinline__ float dpm_x0(const SamplerK &k, const float *c, float x, float m) {
  const float e = (k.param == 2) ? add(mul(c[0], m), mul(c[1], x)) : m;       // noise_pred_fn, model_type v | noise
  return __fdiv_rn(sub(x, mul(c[1], e)), c[0]);                                // data_prediction_fn
}
__device__ __forceinline__ float dpm_update(const float *c, float x, float m0, float m1) {
  const float first = sub(mul(c[2], x), mul(c[3], m0));
  if (c[6] < 1.5f) return first;                                               // dpm_solver_first_update
  return sub(first, mul(c[5], mul(c[4], sub(m0, m1))));                        // multistep second update, 'dpmsolver'
}

template <int VEC>
__global__ void __launch_bounds__(256) dpm_kernel(const SamplerK k) {
  __shared__ float c[B200_SAMPLER_NCOEF];
  if (threadIdx.x < B200_SAMPLER_NCOEF) {
    const int step = k.step_ptr ? *k.step_ptr : k.step_host;
    c[threadIdx.x] = k.coef[(size_t)step * B200_SAMPLER_NCOEF + threadIdx.x];
  }
  __syncthreads();
  const long long nvec = k.total / VEC;
  const long long chw = (long long)k.C * k.HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e0 = i * VEC;
    float xv[VEC], mv[VEC], hv[VEC], x0[VEC], xn[VEC];
    auto ld = [&](const float *p, float *dst) {
      if (VEC == 4) *reinterpret_cast<float4 *>(dst) = *reinterpret_cast<const float4 *>(p + e0);
      else dst[0] = p[e0];
    };
    auto st = [&](float *p, const float *src) {
      if (VEC == 4) *reinterpret_cast<float4 *>(p + e0) = *reinterpret_cast<const float4 *>(src);
      else p[e0] = src[0];
    };
    ld(k.x, xv);
    if (k.phase == 2) ld(k.pred_x0, x0); else ld(k.m, mv);
    if (k.phase != 1) ld(k.hist, hv);
    const float s = (k.phase == 2) ? k.thresh[e0 / chw] : 1.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (k.phase == 2) x0[j] = __fdiv_rn(fminf(fmaxf(x0[j], -s), s), s);
      else x0[j] = dpm_x0(k, c, xv[j], mv[j]);
      if (k.phase != 1) xn[j] = dpm_update(c, xv[j], x0[j], hv[j]);
    }
    if (k.phase == 1) { st(k.pred_x0, x0); continue; }
    st(k.x_prev, xn);
    st(k.hist, x0);
    if (k.pred_x0) st(k.pred_x0, x0);
  }
}

// ---- per-sample quantile of |x| by exact radix select (dynamic thresholding) ------------------------------------
// |x| as uint bits is monotone; 4 passes of 8 bits narrow the k-th smallest key.  One CTA per sample.
__device__ unsigned radix_select(const float *x, long long M, long long k, unsigned *hist /*[256] smem*/, unsigned *sh /*[2]*/) {
  unsigned prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (long long i = threadIdx.x; i < M; i += blockDim.x) {
      const unsigned key = __float_as_uint(x[i]) & 0x7fffffffu;
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long acc = 0;
      unsigned b = 0;
      for (; b < 256; ++b) {
        if (acc + hist[b] > (unsigned long long)k) break;
        acc += hist[b];
      }
      sh[0] = b;
      sh[1] = (unsigned)acc;
    }
    __syncthreads();
    prefix |= sh[0] << shift;
    mask |= 255u << shift;
    k -= sh[1];
    __syncthreads();
  }
  return prefix;
}

__global__ void __launch_bounds__(1024) abs_quantile_kernel(const float *x, long long M, float q, float floor_val, float *out) {
  __shared__ unsigned hist[256];
  __shared__ unsigned sh[2];
  const float *xs = x + (size_t)blockIdx.x * M;
  const float rank = __fmul_rn(q, (float)(M - 1));          // torch.quantile: ranks = q * (n - 1) in the input dtype
  const float lo_f = floorf(rank);
  const long long lo = (long long)lo_f;
  const long long hi = min(lo + 1, M - 1);                   // ceil(rank) (== lo when rank is integral: weight 0)
  const float w = rank - lo_f;
  const float vlo = __uint_as_float(radix_select(xs, M, lo, hist, sh));
  const float vhi = (w > 0.f) ? __uint_as_float(radix_select(xs, M, hi, hist, sh)) : vlo;
  if (threadIdx.x == 0) {
    const float d = vhi - vlo;                               // torch.lerp
    const float v = (w < 0.5f) ? vlo + w * d : vhi - d * (1.f - w);
    out[blockIdx.x] = fmaxf(v, floor_val);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) sampler_kernel(const SamplerK k) {
  __shared__ float c[B200_SAMPLER_NCOEF];
  if (threadIdx.x < B200_SAMPLER_NCOEF) {
    const int step = k.step_ptr ? *k.step_ptr : k.step_host;
    c[threadIdx.x] = k.coef[(size_t)step * B200_SAMPLER_NCOEF + threadIdx.x];
  }
  __syncthreads();
  const long long nvec = k.total / VEC;
  const long long chw = (long long)k.C * k.HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long e0 = i * VEC;
    // model_out is [N, C or 2C, HW]: locate the mean (and variance) planes of this element
    long long mi = e0;
    if (k.learned_range) {
      const long long n = e0 / chw;
      mi = e0 + n * chw;  // n*2*chw + (e0 - n*chw)
    }
    float xv[VEC], mv[VEC], sv[VEC], zv[VEC], xp[VEC], x0[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4 *>(xv) = __ldg(reinterpret_cast<const float4 *>(k.x + e0));
      *reinterpret_cast<float4 *>(mv) = __ldg(reinterpret_cast<const float4 *>(k.m + mi));
      if (k.learned_range && k.mode == B200_SAMPLER_GD_DDPM)
        *reinterpret_cast<float4 *>(sv) = __ldg(reinterpret_cast<const float4 *>(k.m + mi + chw));
      if (k.noise) *reinterpret_cast<float4 *>(zv) = __ldg(reinterpret_cast<const float4 *>(k.noise + e0));
    } else {
      xv[0] = k.x[e0];
      mv[0] = k.m[mi];
      if (k.learned_range && k.mode == B200_SAMPLER_GD_DDPM) sv[0] = k.m[mi + chw];
      if (k.noise) zv[0] = k.noise[e0];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float s = (k.learned_range && k.mode == B200_SAMPLER_GD_DDPM) ? sv[j] : 0.f;
      const float z = k.noise ? zv[j] : 0.f;
      sampler_elem(k, c, xv[j], mv[j], s, z, xp[j], x0[j]);
    }
    if (VEC == 4) {
      *reinterpret_cast<float4 *>(k.x_prev + e0) = *reinterpret_cast<float4 *>(xp);
      if (k.pred_x0) *reinterpret_cast<float4 *>(k.pred_x0 + e0) = *reinterpret_cast<float4 *>(x0);
    } else {
      k.x_prev[e0] = xp[0];
      if (k.pred_x0) k.pred_x0[e0] = x0[0];
    }
  }
}

// out = c1[n] * x (+|-) c2[n] * y with per-sample coefficients: q_sample, predict_start_from_noise, predict_start_from_z_and_v,
// predict_eps_from_z_and_v, get_v, posterior mean (ldm/models/diffusion/ddpm.py:284-311; gaussian_diffusion.py:
// 184-199, 223-245) — the reference's operation order (two multiplies, then one add / subtract; no FMA contraction)
__global__ void __launch_bounds__(256) lincomb2_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                       const float *__restrict__ c1, const float *__restrict__ c2,
                                                       float *__restrict__ out, long long per_sample, long long total,
                                                       int subtract) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / per_sample;
    const float a = mul(c1[n], x[i]), b = mul(c2[n], y[i]);
    out[i] = subtract ? sub(a, b) : add(a, b);
  }
}

__global__ void advance_step_kernel(int *step, int delta) { *step += delta; }

}  // namespace b200

extern "C" int b200_sampler_update(const b200_sampler_desc *d, void *stream) {
  using namespace b200;
  B200_CHECK(d != nullptr, "sampler: null descriptor");
  B200_CHECK(d->mode >= 0 && d->mode <= 4, "sampler: bad mode %d", d->mode);
  B200_CHECK(d->param >= 0 && d->param <= 2, "sampler: bad parameterization %d", d->param);
  B200_CHECK(d->x && d->coef && (d->x_prev || (d->mode == B200_SAMPLER_DPMPP_2M && d->phase == 1)) &&
                 (d->model_out || (d->mode == B200_SAMPLER_DPMPP_2M && d->phase == 2)), "sampler: null tensor");
  if (d->mode == B200_SAMPLER_DPMPP_2M) {
    B200_CHECK(d->phase >= 0 && d->phase <= 2 && d->param != 1 && !d->learned_range, "sampler: bad DPM-Solver++ request");
    B200_CHECK(d->phase == 1 ? d->pred_x0 != nullptr : d->hist != nullptr, "sampler: DPM-Solver++ needs hist / pred_x0");
    B200_CHECK(d->phase != 2 || (d->thresh && d->pred_x0), "sampler: DPM-Solver++ phase 2 needs thresh and pred_x0");
  }
  B200_CHECK(d->N > 0 && d->C > 0 && d->HW > 0, "sampler: empty tensor");
  B200_CHECK(!(d->learned_range && d->mode < 2), "sampler: learned_range is a guided-diffusion mode");
  SamplerK k;
  k.mode = d->mode; k.param = d->param; k.clip = d->clip_denoised; k.learned_range = d->learned_range;
  k.C = d->C; k.HW = d->HW;
  k.total = (long long)d->N * d->C * d->HW;
  k.x = d->x; k.m = d->model_out; k.noise = d->noise; k.coef = d->coef;
  k.step_ptr = d->step_ptr; k.step_host = d->step_host;
  k.x_prev = d->x_prev; k.pred_x0 = d->pred_x0;
  k.hist = d->hist; k.thresh = d->thresh; k.phase = d->phase;
  const bool vec4 = ((long long)d->C * d->HW) % 4 == 0 &&
                    (((uintptr_t)d->x | (uintptr_t)d->model_out | (uintptr_t)d->noise |
                      (uintptr_t)d->x_prev | (uintptr_t)d->pred_x0) & 15) == 0;
  const long long nvec = vec4 ? k.total / 4 : k.total;
  const int blocks = (int)std::min<long long>(cdiv64(nvec, 256), (long long)num_sms() * 8);
  if (d->mode == B200_SAMPLER_DPMPP_2M) {
    const bool v4 = vec4 && (((uintptr_t)d->hist | (uintptr_t)d->thresh) & 15) == 0;
    const long long nv = v4 ? k.total / 4 : k.total;
    const int nb = (int)std::min<long long>(cdiv64(nv, 256), (long long)num_sms() * 8);
    if (v4) dpm_kernel<4><<<nb, 256, 0, (cudaStream_t)stream>>>(k);
    else dpm_kernel<1><<<nb, 256, 0, (cudaStream_t)stream>>>(k);
  } else if (vec4) sampler_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(k);
  else sampler_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(k);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_advance_step(int *step_ptr, int delta, void *stream) {
  using namespace b200;
  B200_CHECK(step_ptr != nullptr, "advance_step: null pointer");
  advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_ptr, delta);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_abs_quantile(const float *x, int N, int64_t M, float q, float floor_val, float *out, void *stream) {
  using namespace b200;
  B200_CHECK(x && out && N > 0 && M > 0 && q >= 0.f && q <= 1.f, "abs_quantile: bad arguments");
  abs_quantile_kernel<<<N, 1024, 0, (cudaStream_t)stream>>>(x, (long long)M, q, floor_val, out);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_lincomb2(const float *x, const float *y, const float *c1, const float *c2, float *out, int N,
                             int64_t per_sample, int subtract, void *stream) {
  using namespace b200;
  B200_CHECK(x && y && c1 && c2 && out && N > 0 && per_sample > 0, "lincomb2: bad arguments");
  const long long total = (long long)N * per_sample;
  const int blocks = (int)std::min<long long>(cdiv64(total, 256), (long long)num_sms() * 8);
  lincomb2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, c1, c2, out, per_sample, total, subtract);
  B200_LAUNCH_CHECK();
  return 0;
}
