// Per-ray kernels (one warp per ray): coarse sampling, hierarchical up-sampling step (density ->
// transmittance scans -> inverse-CDF resampling -> sorted merge), render_core pre/post MLP stages.
//
// Replaces (paths relative to /root/reference/src/models/udf_renderer_blending.py):
//   :705-720  coarse z                       -> coarse_z_kernel
//   :228-353  up_sample_unbias               -> upsample_step_kernel   (mode 0)
//   :920-975  up_sample_no_occ_aware         -> upsample_step_kernel   (mode 1)
//   :69-109   sample_pdf(det=True)           -> inside upsample_step_kernel
//   :355-377  cat_z_vals (cat+sort+gather)   -> 2-way rank merge inside upsample_step_kernel
//   :435-455  dists / mid_z                  -> render_prep_kernel
//   :463-650  render_core after the MLP      -> render_core_fwd_kernel (+ render_reduce_kernel)
//
// These stages are HBM/latency bound (a few hundred bytes per ray-sample) and fp32 element-wise;
// the transmittance products and the CDF are scanned in fp64 -- this reproduces torch's CPU
// cumprod/cumsum (which accumulate fp32 inputs in double) so that, given identical inputs,
// searchsorted picks identical bins.  Compiled with -fmad=false: the reference evaluates every
// a*b+c as two rounded ops.
#include <math.h>

#include "common.cuh"
#include "host.h"

namespace emap {

constexpr int kMaxSamples = 512;   // samples per ray (n_samples + n_importance)
constexpr int kMaxNew = 64;        // new samples per up-sampling step
constexpr int kRayWarps = 4;       // warps (= rays) per block
// floats per per-ray shared-memory array for m samples (a multiple of 32, one spare row: bank spread)
__host__ __device__ inline int ray_stride(int m) { return ((m + 31) & ~31) + 32; }

__device__ __forceinline__ double shfl_up_d(double v, int d) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_up_sync(0xffffffffu, lo, d);
  hi = __shfl_up_sync(0xffffffffu, hi, d);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o; o >>= 1) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, o);
    hi = __shfl_xor_sync(0xffffffffu, hi, o);
    v += __hiloint2double(hi, lo);
  }
  return v;
}

// Scans over a ray's m samples (fp32 in SMEM, accumulated in fp64 like torch's CPU cumprod / cumsum).  BLOCKED: lane l
// owns the C = ceil(m/32) consecutive samples [l C, l C + C): a sequential local pass, ONE 5-level warp scan of the
// lane totals, a second local pass -- C + 5 + C dependent fp64 operations instead of the 5 C of a row-by-row
// shuffle scan (C = 8 at 256 samples; these kernels are latency-bound: one warp per ray, < 1 wave at 4096 rays).
// In-place allowed (a lane reads and writes only its own block).  All 32 lanes participate.

// Exclusive running product: out[i] = float(prod_{t<i} arr[t]).
__device__ __forceinline__ void excl_cumprod(const float* arr, float* out, int m, int lane) {
  const int C = (m + 31) >> 5, i0 = lane * C;
  double loc = 1.0;
  for (int t = 0; t < C; ++t) { const int i = i0 + t; if (i < m) loc *= (double)arr[i]; }
  double inc = loc;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double t = shfl_up_d(inc, d);
    if (lane >= d) inc *= t;
  }
  double run = shfl_up_d(inc, 1);
  if (lane == 0) run = 1.0;
  for (int t = 0; t < C; ++t) {
    const int i = i0 + t;
    if (i < m) { const double v = (double)arr[i]; out[i] = (float)run; run *= v; }
  }
  __syncwarp();
}

// Inclusive running sum shifted by one: out[0] = 0, out[i + 1] = float(sum_{t<=i} arr[t] / total)   (the CDF)
__device__ __forceinline__ void cdf_cumsum(const float* arr, float* out, int m, float total, int lane) {
  const int C = (m + 31) >> 5, i0 = lane * C;
  double loc = 0.0;
  for (int t = 0; t < C; ++t) { const int i = i0 + t; if (i < m) loc += (double)(arr[i] / total); }
  double inc = loc;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double t = shfl_up_d(inc, d);
    if (lane >= d) inc += t;
  }
  double run = shfl_up_d(inc, 1);
  if (lane == 0) { run = 0.0; out[0] = 0.f; }
  for (int t = 0; t < C; ++t) {
    const int i = i0 + t;
    if (i < m) { run += (double)(arr[i] / total); out[i + 1] = (float)run; }
  }
  __syncwarp();
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// udf2logistic(udf, inv_s, gamma=1, abs_cos=1)  (udf_renderer_blending.py:163-170)
__device__ __forceinline__ float udf2logistic(float udf, float inv_s) {
  const float ex = expf(-inv_s * udf);
  const float opl = 1.0f + ex;
  return (1.0f * inv_s) * ex / (opl * opl) * 1.0f;
}

// sdf2alpha (udf_renderer_blending.py:379-416); type 0 = "numerical", 1 = "theorical".
// `ratio` < 0 means cos_anneal_ratio=None.
__device__ __forceinline__ float sdf2alpha(float sdf, float true_cos, float dists, float inv_s,
                                           float ratio, int type) {
  float iter_cos = true_cos;
  if (ratio >= 0.f)
    iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.0f - ratio) + fmaxf(-true_cos, 0.f) * ratio);
  if (type == 0) {
    const float hstep = iter_cos * dists * 0.5f;
    const float prev_cdf = sigmoidf_((sdf - hstep) * inv_s);
    const float next_cdf = sigmoidf_((sdf + hstep) * inv_s);
    const float a = ((prev_cdf - next_cdf) + 1e-5f) / (prev_cdf + 1e-5f);
    return fminf(fmaxf(a, 0.f), 1.f);
  }
  const float raw = fabsf(iter_cos) * inv_s * (1.0f - sigmoidf_(sdf * inv_s));
  return 1.0f - expf(-fmaxf(raw, 0.f) * dists);
}

// --------------------------------------------------------------------------------------------
__global__ void coarse_z_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                int near_stride, const float* __restrict__ lin,
                                const float* __restrict__ t_rand, int B, int n, float* __restrict__ z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * n) return;
  const int b = (int)(idx / n), i = (int)(idx % n);
  const float nr = near[(size_t)b * near_stride], fr = far[(size_t)b * near_stride];
  float v = nr + (fr - nr) * lin[i];                       // :707
  if (t_rand) v = v + t_rand[b] * 2.0f / (float)n;         // :720
  z[idx] = v;
}

// --------------------------------------------------------------------------------------------
struct UpsampleArgs {
  const float* rays_o; const float* rays_d;
  const float* z_in; const float* udf_in; int n;           // current samples (sorted)
  const float* z_add; const float* udf_add; int ka;        // pending new samples to merge first (sorted)
  float* z_out; float* udf_out;                            // merged [B, n+ka] (written iff ka>0)
  const float* u;                                          // [k] quantiles linspace(.5/k, 1-.5/k, k)
  float* z_new; long long* inds; float* w_out;             // [B,k] sorted, [B,k] optional, [B,m-1] optional
  const float* sample_dist;                                // device scalar
  const float* gamma_ptr;                                  // optional device scalar overriding `gamma`
  int B, k;
  float inv_s, beta, gamma;
  int mode;        // 0 unbias (occlusion aware), 1 no_occ_aware, 2 weights given (sample_pdf only)
  int alpha_type;  // 0 numerical, 1 theorical
  int* status;     // optional device status word: EMAP_STATUS_NAN_SAMPLES when a new sample is NaN (:102, :346)
};

__global__ void __launch_bounds__(kRayWarps * 32) upsample_step_kernel(const UpsampleArgs a) {
  // per-ray arrays in DYNAMIC shared memory sized by the actual sample count (ray_stride(n + ka) floats each): with
  // static [kMaxSamples] arrays a block took 36 KiB -> 6 blocks per SM, and 4096 rays (1024 blocks) needed 1.15 waves
  extern __shared__ float s_dyn[];
  __shared__ float s_new[kRayWarps][kMaxNew];
  __shared__ float s_add[kRayWarps][2 * kMaxNew];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kRayWarps + wib;
  if (ray >= a.B) return;
  const int n = a.n, ka = a.ka, m = n + ka;
  const int rs = ray_stride(m);
  float* zs = s_dyn + (size_t)(wib * 4 + 0) * rs;
  float* us = s_dyn + (size_t)(wib * 4 + 1) * rs;
  float* sa = s_dyn + (size_t)(wib * 4 + 2) * rs;     // scratch: vp terms / alpha / cdf
  float* sb = s_dyn + (size_t)(wib * 4 + 3) * rs;     // scratch: vis_prob / weights
  float* zn = s_new[wib];

  // ---- (a12) merge pending samples: z sorted union, udf permuted alike
  if (ka > 0) {
    float* za = s_add[wib]; float* ua = za + kMaxNew;
    for (int j = lane; j < ka; j += 32) { za[j] = a.z_add[(size_t)ray * ka + j]; ua[j] = a.udf_add ? a.udf_add[(size_t)ray * ka + j] : 0.f; }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const float zi = a.z_in[(size_t)ray * n + i];
      int lo = 0, hi = ka;                 // # add elements strictly below zi
      while (lo < hi) { int mid = (lo + hi) >> 1; if (za[mid] < zi) lo = mid + 1; else hi = mid; }
      zs[i + lo] = zi;
      us[i + lo] = a.udf_in ? a.udf_in[(size_t)ray * n + i] : 0.f;
    }
    const float* zin = a.z_in + (size_t)ray * n;
    for (int j = lane; j < ka; j += 32) {
      const float zj = za[j];
      int lo = 0, hi = n;                  // # old elements <= zj  (old first on ties)
      while (lo < hi) { int mid = (lo + hi) >> 1; if (zin[mid] <= zj) lo = mid + 1; else hi = mid; }
      zs[j + lo] = zj;
      us[j + lo] = ua[j];
    }
    __syncwarp();
    for (int i = lane; i < m; i += 32) {
      a.z_out[(size_t)ray * m + i] = zs[i];
      if (a.udf_out) a.udf_out[(size_t)ray * m + i] = us[i];
    }
  } else if (a.mode == 2) {
    for (int i = lane; i < n; i += 32) zs[i] = a.z_in[(size_t)ray * n + i];
  } else {
    for (int i = lane; i < n; i += 32) { zs[i] = a.z_in[(size_t)ray * n + i]; us[i] = a.udf_in[(size_t)ray * n + i]; }
  }
  __syncwarp();
  const int k = a.k;
  if (k <= 0) return;

  const float ox = a.rays_o[ray * 3 + 0], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
  const float dx = a.rays_d[ray * 3 + 0], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
  const float sample_dist = *a.sample_dist;
  const float inv_s = a.inv_s, beta = a.beta, gamma = a.gamma_ptr ? *a.gamma_ptr : a.gamma;

  // ---- weights[0..m-1)
  if (a.mode == 0) {
    // vp term per sample i in [0,m):  clip(1 - alpha_occ + vis_mask, 0, 1) + 1e-7     (:293-319)
    for (int i = lane; i < m; i += 32) {
      const float dist_raw = (i + 1 < m) ? (zs[i + 1] - zs[i]) : sample_dist;
      const float raw_occ = udf2logistic(us[i], beta);
      const float alpha_occ = 1.0f - expf(-fmaxf(raw_occ, 0.f) * gamma * dist_raw);
      float vis_mask = 1.0f;
      if (i > 0) {
        const float tc = (us[i] - us[i - 1]) / (zs[i] - zs[i - 1] + 1e-5f);
        vis_mask = (tc < 0.05f) ? 1.0f : 0.0f;
      }
      sa[i] = fminf(fmaxf((1.0f - alpha_occ) + vis_mask, 0.f), 1.f) + 1e-7f;
    }
    __syncwarp();
    excl_cumprod(sa, sb, m, lane);        // sb[i] = vis_prob_i
    // alpha per interval i in [0,m-1)                                                  (:279-332)
    for (int i = lane; i < m - 1; i += 32) {
      const float z0 = zs[i], z1 = zs[i + 1], u0 = us[i], u1 = us[i + 1];
      const float r0 = sqrtf((ox + dx * z0) * (ox + dx * z0) + (oy + dy * z0) * (oy + dy * z0) + (oz + dz * z0) * (oz + dz * z0));
      const float r1 = sqrtf((ox + dx * z1) * (ox + dx * z1) + (oy + dy * z1) * (oy + dy * z1) + (oz + dz * z1) * (oz + dz * z1));
      const float inside = (r0 < 1.0f || r1 < 1.0f) ? 1.0f : 0.0f;
      const float tc = (u1 - u0) / (z1 - z0 + 1e-5f);
      float cosv = -fabsf(tc);
      float prevc = 0.f;
      if (i > 0) prevc = -fabsf((u0 - us[i - 1]) / (z0 - zs[i - 1] + 1e-5f));
      cosv = fminf(prevc, cosv);
      cosv = fminf(fmaxf(cosv, -1e3f), 0.0f) * inside;
      const float mid_udf = (u0 + u1) * 0.5f;
      const float dists = z1 - z0;
      const float ap = sdf2alpha(mid_udf, cosv, dists, inv_s, -1.f, a.alpha_type);
      const float am = sdf2alpha(mid_udf * -1.f, cosv, dists, inv_s, -1.f, a.alpha_type);
      const float sp = sb[i];
      const float alpha = ap * sp + am * (1.0f - sp);
      sa[i] = alpha;
    }
    __syncwarp();
    // weights = alpha * exclusive cumprod(1 - alpha + 1e-7)                             (:334-343)
    for (int i = lane; i < m - 1; i += 32) sb[i] = (1.0f - sa[i]) + 1e-7f;
    __syncwarp();
    excl_cumprod(sb, sb, m - 1, lane);
    for (int i = lane; i < m - 1; i += 32) sb[i] = sa[i] * sb[i];
  } else if (a.mode == 2) {
    // sample_pdf in isolation: udf_in holds the weights [B, m-1] as given
    for (int i = lane; i < m - 1; i += 32) sb[i] = a.udf_in[(size_t)ray * (m - 1) + i];
  } else {
    // no_occ_aware: weights = alpha_occ[:, :-1]                                         (:945-967)
    for (int i = lane; i < m - 1; i += 32) {
      const float dist = zs[i + 1] - zs[i];
      const float raw_occ = udf2logistic(us[i], beta);
      sb[i] = 1.0f - expf(-fmaxf(raw_occ, 0.f) * gamma * dist);
    }
  }
  __syncwarp();
  if (a.w_out)
    for (int i = lane; i < m - 1; i += 32) a.w_out[(size_t)ray * (m - 1) + i] = sb[i];

  // ---- (a11b) sample_pdf(z, weights, k, det=True)                                      (:69-109)
  double part = 0.0;
  for (int i = lane; i < m - 1; i += 32) { sb[i] = sb[i] + 1e-5f; part += (double)sb[i]; }
  const float total = (float)warp_sum_d(part);
  __syncwarp();
  cdf_cumsum(sb, sa, m - 1, total, lane);  // cdf[0] = 0; cdf[i+1] = float(cumsum_fp64(pdf)[i]);  stored in sa[0..m)
  for (int j = lane; j < k; j += 32) {
    const float uj = a.u[j];
    int lo = 0, hi = m;                    // searchsorted(cdf, u, right=True): # cdf entries <= u
    while (lo < hi) { int mid = (lo + hi) >> 1; if (sa[mid] <= uj) lo = mid + 1; else hi = mid; }
    const int ind = lo;
    const int below = max(ind - 1, 0), above = min(m - 1, ind);
    const float c0 = sa[below], c1 = sa[above], b0 = zs[below], b1 = zs[above];
    float denom = c1 - c0;
    if (denom < 1e-5f) denom = 1.0f;
    const float t = (uj - c0) / denom;
    zn[j] = b0 + t * (b1 - b0);
    if (a.inds) a.inds[(size_t)ray * k + j] = (long long)ind;
  }
  __syncwarp();
  // rank-sort the k new samples (they are ascending up to rounding at bin boundaries)
  for (int j = lane; j < k; j += 32) {
    const float v = zn[j];
    int rank = 0;
    for (int t = 0; t < k; ++t) { const float o = zn[t]; rank += (o < v || (o == v && t < j)) ? 1 : 0; }
    a.z_new[(size_t)ray * k + rank] = v;
    // the reference's NaN guard on the new samples (pdb.set_trace() at :102-107 and :346-351): a device flag
    // the host polls once per step.  (A NaN ranks 0 like every NaN here; the flag is what matters.)
    if (a.status && v != v) atomicOr(a.status, EMAP_STATUS_NAN_SAMPLES);
  }
}

// --------------------------------------------------------------------------------------------
// render_core, before the MLP: dists, mid_z                                      (:435-446)
__global__ void render_prep_kernel(const float* __restrict__ z, const float* __restrict__ sample_dist,
                                   int B, int n, float* __restrict__ dists, float* __restrict__ mid_z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * n) return;
  const int i = (int)(idx % n);
  const float z0 = z[idx];
  const float d = (i + 1 < n) ? (z[idx + 1] - z0) : *sample_dist;
  dists[idx] = d;
  mid_z[idx] = z0 + d * 0.5f;
}

struct CoreArgs {
  const float* rays_o; const float* rays_d;
  const float* mid_z; const float* dists;       // [B,n]
  const float* udf; const float* grad;          // [B*n], [B*n,3]  (MLP outputs at the mid points)
  const float* scalars;                         // device: [inv_s, beta, gamma]
  int B, n;
  float cos_anneal_ratio;                       // < 0 : None
  float flip_saturation, near_surface, sparse_scale;
  int use_unbias, use_norm_grad, alpha_type;
  // outputs
  float* weights; float* alpha; float* grad_flip; float* inside_sphere; float* grad_mag;   // per sample
  float* edge; float* depth; float* normals;                                               // per ray
  double* partials;                             // [B,5]: sum(relax*gerr), sum(relax), sum(near*gerr), sum(near), sum exp(-s*udf)
};

__global__ void __launch_bounds__(kRayWarps * 32) render_core_fwd_kernel(const CoreArgs a) {
  extern __shared__ float s_dyn[];                  // 3 arrays of ray_stride(n) floats per ray (see upsample_step_kernel)
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kRayWarps + wib;
  if (ray >= a.B) return;
  const int n = a.n;
  const int rs = ray_stride(n);
  float* sa = s_dyn + (size_t)(wib * 3 + 0) * rs;
  float* sb = s_dyn + (size_t)(wib * 3 + 1) * rs;
  float* stc = s_dyn + (size_t)(wib * 3 + 2) * rs;
  const size_t base = (size_t)ray * n;
  const float ox = a.rays_o[ray * 3 + 0], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
  const float dx = a.rays_d[ray * 3 + 0], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
  const float inv_s = a.scalars[0], beta = a.scalars[1], gamma = a.scalars[2];

  double p_rg = 0.0, p_r = 0.0, p_ng = 0.0, p_n = 0.0, p_sp = 0.0;
  // pass 1: per-sample quantities that need no scan; true_cos -> stc
  for (int i = lane; i < n; i += 32) {
    const float gx = a.grad[(base + i) * 3 + 0], gy = a.grad[(base + i) * 3 + 1], gz = a.grad[(base + i) * 3 + 2];
    const float udf = a.udf[base + i];
    const float mz = a.mid_z[base + i];
    const float mag = sqrtf(gx * gx + gy * gy + gz * gz);                         // :463
    const float inv = mag + 1e-5f;
    const float nx = gx / inv, ny = gy / inv, nz = gz / inv;
    const float cosn = dx * nx + dy * ny + dz * nz;                                // :485
    float tc = a.use_norm_grad ? cosn : (dx * gx + dy * gy + dz * gz);            // :479-482
    stc[i] = tc;
    float flip = (cosn > 0.f) ? -1.f : ((cosn < 0.f) ? 1.f : 1.f);                 // :486-489
    if (!a.use_unbias) flip = 1.f;
    a.grad_flip[(base + i) * 3 + 0] = flip * gx;
    a.grad_flip[(base + i) * 3 + 1] = flip * gy;
    a.grad_flip[(base + i) * 3 + 2] = flip * gz;
    a.grad_mag[base + i] = mag;
    const float px = ox + dx * mz, py = oy + dy * mz, pz = oz + dz * mz;
    const float pn = sqrtf(px * px + py * py + pz * pz);                           // :563
    a.inside_sphere[base + i] = (pn < 2.0f) ? 1.f : 0.f;                           // :568
    const float relax = (pn < 2.4f) ? 1.f : 0.f;                                   // :569
    const float near = (udf < a.near_surface) ? 1.f : 0.f;                         // :570
    const float gerr = (mag - 1.0f) * (mag - 1.0f);                                // :612-617
    p_rg += (double)(relax * gerr); p_r += (double)relax;
    p_ng += (double)(near * gerr);  p_n += (double)near;
    p_sp += (double)expf(-a.sparse_scale * udf);                                   // :642
  }
  __syncwarp();
  if (a.use_unbias) {
    for (int i = lane; i < n; i += 32) {
      const float udf = a.udf[base + i], dist = a.dists[base + i];
      const float raw_occ = udf2logistic(udf, beta);                               // :492
      const float alpha_occ = 1.0f - expf(-fmaxf(raw_occ, 0.f) * gamma * dist);    // :497
      const float vm = (i + 1 < n) ? ((stc[i + 1] < 0.01f) ? 1.f : 0.f) : 1.f;      // :500-509
      sa[i] = fminf(fmaxf((1.0f - alpha_occ) + a.flip_saturation * vm, 0.f), 1.f) + 1e-7f;
    }
    __syncwarp();
    excl_cumprod(sa, sb, n, lane);                                                 // :511-523
    for (int i = lane; i < n; i += 32) {
      const float vp = fminf(fmaxf(sb[i], 0.f), 1.f);                              // :528
      const float udf = a.udf[base + i], dist = a.dists[base + i];
      const float nac = -1.f * fabsf(stc[i]);
      const float ap = sdf2alpha(udf, nac, dist, inv_s, a.cos_anneal_ratio, a.alpha_type);
      const float am = sdf2alpha(-udf, nac, dist, inv_s, a.cos_anneal_ratio, a.alpha_type);
      sa[i] = ap * vp + am * (1.0f - vp);                                          // :545
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const float udf = a.udf[base + i], dist = a.dists[base + i];
      const float raw_occ = udf2logistic(udf, beta);
      sa[i] = 1.0f - expf(-fmaxf(raw_occ, 0.f) * gamma * dist);                    // :553-559
    }
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) sb[i] = (1.0f - sa[i]) + 1e-7f;
  __syncwarp();
  excl_cumprod(sb, sb, n, lane);                                                   // :593-602
  double e = 0.0, dep = 0.0, n0 = 0.0, n1 = 0.0, n2 = 0.0;
  for (int i = lane; i < n; i += 32) {
    const float alpha = sa[i];
    const float w = alpha * sb[i];
    a.weights[base + i] = w;
    if (a.alpha) a.alpha[base + i] = alpha;
    e += (double)w;
    dep += (double)(a.mid_z[base + i] * w);
    n0 += (double)(a.grad_flip[(base + i) * 3 + 0] * w);
    n1 += (double)(a.grad_flip[(base + i) * 3 + 1] * w);
    n2 += (double)(a.grad_flip[(base + i) * 3 + 2] * w);
  }
  e = warp_sum_d(e); dep = warp_sum_d(dep); n0 = warp_sum_d(n0); n1 = warp_sum_d(n1); n2 = warp_sum_d(n2);
  p_rg = warp_sum_d(p_rg); p_r = warp_sum_d(p_r); p_ng = warp_sum_d(p_ng); p_n = warp_sum_d(p_n);
  p_sp = warp_sum_d(p_sp);
  if (lane == 0) {
    a.edge[ray] = (float)e;                                                        // :604-606 (sampled_edge == 1)
    a.depth[ray] = (float)dep;                                                     // :607
    a.normals[ray * 3 + 0] = (float)n0; a.normals[ray * 3 + 1] = (float)n1; a.normals[ray * 3 + 2] = (float)n2;
    double* pp = a.partials + (size_t)ray * 5;
    pp[0] = p_rg; pp[1] = p_r; pp[2] = p_ng; pp[3] = p_n; pp[4] = p_sp;
  }
}

// Deterministic reduction of the per-ray partial sums -> [gradient_error, gradient_error_near_surface,
// sparse_error, sum_relax, sum_near]                                             (:618-625, :642-644)
__global__ void render_reduce_kernel(const double* __restrict__ partials, int B, float* __restrict__ out,
                                     int* __restrict__ status) {
  __shared__ double sh[5][32];
  double acc[5] = {0, 0, 0, 0, 0};
  for (int r = threadIdx.x; r < B; r += blockDim.x)
    for (int c = 0; c < 5; ++c) acc[c] += partials[(size_t)r * 5 + c];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = 0; c < 5; ++c) {
    acc[c] = warp_sum_d(acc[c]);
    if (lane == 0) sh[c][w] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5] = {0, 0, 0, 0, 0};
    for (int c = 0; c < 5; ++c) for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t[c] += sh[c][i];
    const float srg = (float)t[0], sr = (float)t[1], sng = (float)t[2], sn = (float)t[3];
    out[0] = srg / (sr + 1e-5f);
    out[1] = sng / (sn + 1e-5f);
    out[2] = (float)(t[4] / (double)B);
    out[3] = sr;
    out[4] = sn;
    // the reference's guard on gradient_error (:632-633)
    if (status && out[0] != out[0]) atomicOr(status, EMAP_STATUS_NAN_EIKONAL);
  }
}


// --------------------------------------------------------------------------------------------
// K3 backward: cotangents of (weights, edge, depth, normals, gradient_error[_near_surface],
// sparse_error) -> cotangents of (udf, grad, inv_s, beta, gamma).  One warp per ray; the forward
// quantities are recomputed (cheaper than storing them); the two transmittance products are
// differentiated with exclusive suffix sums:  d/dm_t prod_{t<i} m_t = (prod)/m_t.
constexpr int kBwdWarps = 2;

// out[t] = float( sum_{i>t} arr[i] )  in fp64 (blocked like excl_cumprod, from the right)
__device__ __forceinline__ void excl_suffix_sum(const float* arr, float* out, int m, int lane) {
  const int C = (m + 31) >> 5, i0 = lane * C;
  double loc = 0.0;
  for (int t = C - 1; t >= 0; --t) { const int i = i0 + t; if (i < m) loc += (double)arr[i]; }
  double inc = loc;                       // inclusive suffix over the lanes: lane j gets sum_{l>=j}
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int lo = __double2loint(inc), hi = __double2hiint(inc);
    lo = __shfl_down_sync(0xffffffffu, lo, d);
    hi = __shfl_down_sync(0xffffffffu, hi, d);
    if (lane + d < 32) inc += __hiloint2double(hi, lo);
  }
  double run = inc - loc;                 // everything to the right of my block
  for (int t = C - 1; t >= 0; --t) {
    const int i = i0 + t;
    if (i < m) { const double v = (double)arr[i]; out[i] = (float)run; run += v; }
  }
  __syncwarp();
}

struct AlphaGrad { float d_sdf, d_itercos, d_invs; };

// backward of sdf2alpha w.r.t. (sdf, iter_cos, inv_s) for upstream cotangent `da`
__device__ __forceinline__ AlphaGrad sdf2alpha_bwd(float sdf, float iter_cos, float dists, float inv_s,
                                                  int type, float da) {
  AlphaGrad g = {0.f, 0.f, 0.f};
  if (da == 0.f) return g;
  if (type == 0) {
    const float hstep = iter_cos * dists * 0.5f;
    const float prv = sdf - hstep, nxt = sdf + hstep;
    const float pc = sigmoidf_(prv * inv_s), nc = sigmoidf_(nxt * inv_s);
    const float den = pc + 1e-5f;
    const float a_raw = ((pc - nc) + 1e-5f) / den;
    if (a_raw < 0.f || a_raw > 1.f) return g;
    const float d_pc = da * nc / (den * den);
    const float d_nc = -da / den;
    const float sp = pc * (1.f - pc), sn = nc * (1.f - nc);
    const float d_prv = d_pc * inv_s * sp, d_nxt = d_nc * inv_s * sn;
    g.d_invs = d_pc * prv * sp + d_nc * nxt * sn;
    g.d_sdf = d_prv + d_nxt;
    g.d_itercos = (d_nxt - d_prv) * dists * 0.5f;
  } else {
    const float sg = sigmoidf_(sdf * inv_s);
    const float aic = fabsf(iter_cos);
    const float raw = aic * inv_s * (1.0f - sg);
    if (raw <= 0.f) return g;
    const float d_raw = da * dists * expf(-raw * dists);
    g.d_invs = d_raw * (aic * (1.0f - sg) - aic * inv_s * sg * (1.f - sg) * sdf);
    g.d_sdf = -d_raw * aic * inv_s * inv_s * sg * (1.f - sg);
    const float sgn = (iter_cos > 0.f) ? 1.f : ((iter_cos < 0.f) ? -1.f : 0.f);
    g.d_itercos = d_raw * inv_s * (1.0f - sg) * sgn;
  }
  return g;
}

struct CoreBwdArgs {
  const float* rays_o; const float* rays_d; const float* mid_z; const float* dists;
  const float* udf; const float* grad; const float* scalars; const float* reduced;
  int B, n;
  float cos_anneal_ratio, flip_saturation, near_surface, sparse_scale;
  int use_unbias, use_norm_grad, alpha_type;
  const float* d_w; const float* d_edge; const float* d_depth; const float* d_normals;   // may be NULL
  const float* d_gerr; const float* d_gerr_ns; const float* d_sparse;                    // device scalars or NULL
  float* d_udf; float* d_grad; double* partials;   // partials [B,3]: d_inv_s, d_beta, d_gamma
};

__global__ void __launch_bounds__(kBwdWarps * 32) render_core_bwd_kernel(const CoreBwdArgs a) {
  // 7 arrays of ray_stride(n) floats per ray, dynamic (static [kMaxSamples] arrays: 29 KiB per 2-warp block ->
  // 14 resident warps per SM and 2 waves at 4096 rays)
  extern __shared__ float s_dyn[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kBwdWarps + wib;
  if (ray >= a.B) return;
  const int n = a.n;
  const int rs = ray_stride(n);
  float* tcs = s_dyn + (size_t)(wib * 7 + 0) * rs; float* vt = s_dyn + (size_t)(wib * 7 + 1) * rs;
  float* V = s_dyn + (size_t)(wib * 7 + 2) * rs;   float* al = s_dyn + (size_t)(wib * 7 + 3) * rs;
  float* T = s_dyn + (size_t)(wib * 7 + 4) * rs;   float* sx = s_dyn + (size_t)(wib * 7 + 5) * rs;
  float* sy = s_dyn + (size_t)(wib * 7 + 6) * rs;
  const size_t base = (size_t)ray * n;
  const float ox = a.rays_o[ray * 3 + 0], oy = a.rays_o[ray * 3 + 1], oz = a.rays_o[ray * 3 + 2];
  const float dx = a.rays_d[ray * 3 + 0], dy = a.rays_d[ray * 3 + 1], dz = a.rays_d[ray * 3 + 2];
  const float inv_s = a.scalars[0], beta = a.scalars[1], gamma = a.scalars[2];
  const float ratio = a.cos_anneal_ratio;
  const float de = a.d_edge ? a.d_edge[ray] : 0.f;
  const float dd = a.d_depth ? a.d_depth[ray] : 0.f;
  const float dn0 = a.d_normals ? a.d_normals[ray * 3 + 0] : 0.f;
  const float dn1 = a.d_normals ? a.d_normals[ray * 3 + 1] : 0.f;
  const float dn2 = a.d_normals ? a.d_normals[ray * 3 + 2] : 0.f;
  const float dge = a.d_gerr ? a.d_gerr[0] / (a.reduced[3] + 1e-5f) : 0.f;
  const float dgn = a.d_gerr_ns ? a.d_gerr_ns[0] / (a.reduced[4] + 1e-5f) : 0.f;
  const float dsp = a.d_sparse ? a.d_sparse[0] / (float)a.B : 0.f;

  // ---- recompute forward
  for (int i = lane; i < n; i += 32) {
    const float gx = a.grad[(base + i) * 3 + 0], gy = a.grad[(base + i) * 3 + 1], gz = a.grad[(base + i) * 3 + 2];
    float tc = dx * gx + dy * gy + dz * gz;
    if (a.use_norm_grad) { const float inv = sqrtf(gx * gx + gy * gy + gz * gz) + 1e-5f; tc = dx * (gx / inv) + dy * (gy / inv) + dz * (gz / inv); }
    tcs[i] = tc;
  }
  __syncwarp();
  if (a.use_unbias) {
    for (int i = lane; i < n; i += 32) {
      const float udf = a.udf[base + i], dist = a.dists[base + i];
      const float ao = 1.0f - expf(-fmaxf(udf2logistic(udf, beta), 0.f) * gamma * dist);
      const float vm = (i + 1 < n) ? ((tcs[i + 1] < 0.01f) ? 1.f : 0.f) : 1.f;
      vt[i] = fminf(fmaxf((1.0f - ao) + a.flip_saturation * vm, 0.f), 1.f) + 1e-7f;
    }
    __syncwarp();
    excl_cumprod(vt, V, n, lane);
    for (int i = lane; i < n; i += 32) {
      const float vp = fminf(fmaxf(V[i], 0.f), 1.f);
      const float udf = a.udf[base + i], dist = a.dists[base + i];
      const float nac = -1.f * fabsf(tcs[i]);
      const float ap = sdf2alpha(udf, nac, dist, inv_s, ratio, a.alpha_type);
      const float am = sdf2alpha(-udf, nac, dist, inv_s, ratio, a.alpha_type);
      al[i] = ap * vp + am * (1.0f - vp);
    }
  } else {
    for (int i = lane; i < n; i += 32) {
      const float udf = a.udf[base + i], dist = a.dists[base + i];
      al[i] = 1.0f - expf(-fmaxf(udf2logistic(udf, beta), 0.f) * gamma * dist);
    }
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) sx[i] = (1.0f - al[i]) + 1e-7f;     // m_i
  __syncwarp();
  excl_cumprod(sx, T, n, lane);

  // ---- d alpha:  Wbar_i T_i  -  (sum_{j>i} Wbar_j alpha_j T_j) / m_i
  for (int i = lane; i < n; i += 32) {
    const float gx = a.grad[(base + i) * 3 + 0], gy = a.grad[(base + i) * 3 + 1], gz = a.grad[(base + i) * 3 + 2];
    float flip = 1.f;
    if (a.use_unbias) {
      const float inv = sqrtf(gx * gx + gy * gy + gz * gz) + 1e-5f;
      const float cosn = dx * (gx / inv) + dy * (gy / inv) + dz * (gz / inv);
      flip = (cosn > 0.f) ? -1.f : 1.f;
    }
    const float wbar = (a.d_w ? a.d_w[base + i] : 0.f) + de + a.mid_z[base + i] * dd +
                       flip * (gx * dn0 + gy * dn1 + gz * dn2);
    sy[i] = wbar;                              // keep Wbar
    V[i] = V[i];                               // (V stays)
    vt[i] = vt[i];
    // product term for the suffix sum
    al[i] = al[i];
  }
  __syncwarp();
  // suffix over Wbar_j * alpha_j * T_j  -> reuse sx as input (m_i is recomputed from alpha below)
  for (int i = lane; i < n; i += 32) sx[i] = sy[i] * al[i] * T[i];
  __syncwarp();
  excl_suffix_sum(sx, sx, n, lane);            // sx[i] = R_i
  double p_s = 0.0, p_b = 0.0, p_g = 0.0;
  // d_alpha into sy
  for (int i = lane; i < n; i += 32) {
    const float m_i = (1.0f - al[i]) + 1e-7f;
    sy[i] = sy[i] * T[i] - sx[i] / m_i;
  }
  __syncwarp();

  if (a.use_unbias) {
    // d vp_i = d_alpha_i (ap - am);  through clip and the vis product
    for (int i = lane; i < n; i += 32) {
      const float udf = a.udf[base + i], dist = a.dists[base + i];
      const float nac = -1.f * fabsf(tcs[i]);
      const float ap = sdf2alpha(udf, nac, dist, inv_s, ratio, a.alpha_type);
      const float am = sdf2alpha(-udf, nac, dist, inv_s, ratio, a.alpha_type);
      const float Vi = V[i];
      const float dV = (Vi >= 0.f && Vi <= 1.f) ? sy[i] * (ap - am) : 0.f;
      sx[i] = dV * Vi;
    }
    __syncwarp();
    excl_suffix_sum(sx, sx, n, lane);          // sx[t] = sum_{i>t} dV_i V_i
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const float udf = a.udf[base + i], dist = a.dists[base + i], mz = a.mid_z[base + i];
    const float gx = a.grad[(base + i) * 3 + 0], gy = a.grad[(base + i) * 3 + 1], gz = a.grad[(base + i) * 3 + 2];
    const float mag = sqrtf(gx * gx + gy * gy + gz * gz);
    float du = 0.f, dgx = 0.f, dgy = 0.f, dgz = 0.f;
    const float dal = sy[i];
    // logistic density pieces
    const float ex = expf(-beta * udf);
    const float opl = 1.0f + ex;
    const float raw = beta * ex / (opl * opl);
    const float E = expf(-raw * gamma * dist);
    const float fprime = (1.0f - ex) / (opl * opl * opl);          // d/dy [y/(1+y)^2]
    const float draw_du = -beta * beta * ex * fprime;
    const float draw_db = ex / (opl * opl) - beta * udf * ex * fprime;
    float d_ao = 0.f;
    float d_tc = 0.f;
    if (a.use_unbias) {
      const float q = (1.0f - (1.0f - E)) + a.flip_saturation * ((i + 1 < n) ? ((tcs[i + 1] < 0.01f) ? 1.f : 0.f) : 1.f);
      const float d_vt = sx[i] / vt[i];
      const float d_q = (q >= 0.f && q <= 1.f) ? d_vt : 0.f;
      d_ao = -d_q;
      const float vp = fminf(fmaxf(V[i], 0.f), 1.f);
      const float tc = tcs[i];
      const float nac = -1.f * fabsf(tc);
      float iter_cos = nac, dic_dnac = 1.f;
      if (ratio >= 0.f) {
        const float r1 = -nac * 0.5f + 0.5f, r2 = -nac;
        iter_cos = -(fmaxf(r1, 0.f) * (1.0f - ratio) + fmaxf(r2, 0.f) * ratio);
        dic_dnac = ((r1 > 0.f) ? 0.5f * (1.0f - ratio) : 0.f) + ((r2 > 0.f) ? ratio : 0.f);
      }
      const AlphaGrad gp = sdf2alpha_bwd(udf, iter_cos, dist, inv_s, a.alpha_type, dal * vp);
      const AlphaGrad gm = sdf2alpha_bwd(-udf, iter_cos, dist, inv_s, a.alpha_type, dal * (1.0f - vp));
      du += gp.d_sdf - gm.d_sdf;
      p_s += (double)(gp.d_invs + gm.d_invs);
      const float d_nac = (gp.d_itercos + gm.d_itercos) * dic_dnac;
      const float sgn = (tc > 0.f) ? 1.f : ((tc < 0.f) ? -1.f : 0.f);
      d_tc = -d_nac * sgn;
    } else {
      d_ao = dal;
    }
    // ao = 1 - exp(-raw*gamma*dist)
    const float d_raw = d_ao * gamma * dist * E;
    p_g += (double)(d_ao * raw * dist * E);
    du += d_raw * draw_du;
    p_b += (double)(d_raw * draw_db);
    // true_cos -> grad
    if (d_tc != 0.f) {
      if (a.use_norm_grad) {
        const float inv = mag + 1e-5f;
        const float dot = dx * gx + dy * gy + dz * gz;
        const float k = (mag > 0.f) ? dot / (mag * inv * inv) : 0.f;
        dgx += d_tc * (dx / inv - k * gx); dgy += d_tc * (dy / inv - k * gy); dgz += d_tc * (dz / inv - k * gz);
      } else {
        dgx += d_tc * dx; dgy += d_tc * dy; dgz += d_tc * dz;
      }
    }
    // normals = sum flip*g*w
    {
      float flip = 1.f;
      if (a.use_unbias) {
        const float inv = mag + 1e-5f;
        const float cosn = dx * (gx / inv) + dy * (gy / inv) + dz * (gz / inv);
        flip = (cosn > 0.f) ? -1.f : 1.f;
      }
      const float w = al[i] * T[i];
      dgx += flip * w * dn0; dgy += flip * w * dn1; dgz += flip * w * dn2;
    }
    // eikonal terms
    {
      const float px = ox + dx * mz, py = oy + dy * mz, pz = oz + dz * mz;
      const float pn = sqrtf(px * px + py * py + pz * pz);
      const float relax = (pn < 2.4f) ? 1.f : 0.f;
      const float near = (udf < a.near_surface) ? 1.f : 0.f;
      const float coef = (dge * relax + dgn * near) * 2.0f * (mag - 1.0f);
      if (mag > 0.f) { dgx += coef * gx / mag; dgy += coef * gy / mag; dgz += coef * gz / mag; }
    }
    if (dsp != 0.f) du += dsp * (-a.sparse_scale) * expf(-a.sparse_scale * udf);
    a.d_udf[base + i] = du;
    a.d_grad[(base + i) * 3 + 0] = dgx; a.d_grad[(base + i) * 3 + 1] = dgy; a.d_grad[(base + i) * 3 + 2] = dgz;
  }
  p_s = warp_sum_d(p_s); p_b = warp_sum_d(p_b); p_g = warp_sum_d(p_g);
  if (lane == 0) { a.partials[(size_t)ray * 3 + 0] = p_s; a.partials[(size_t)ray * 3 + 1] = p_b; a.partials[(size_t)ray * 3 + 2] = p_g; }
}

__global__ void scalar_reduce_kernel(const double* __restrict__ partials, int B, int ncol, float* __restrict__ out) {
  __shared__ double sh[8][32];
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = threadIdx.x; r < B; r += blockDim.x)
    for (int c = 0; c < ncol; ++c) acc[c] += partials[(size_t)r * ncol + c];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = 0; c < ncol; ++c) { acc[c] = warp_sum_d(acc[c]); if (lane == 0) sh[c][w] = acc[c]; }
  __syncthreads();
  if (threadIdx.x < ncol) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[threadIdx.x][i];
    out[threadIdx.x] = (float)t;
  }
}

}  // namespace emap

using namespace emap;

extern "C" int emap_coarse_z(const float* near, const float* far, int32_t near_is_per_ray,
                             const float* lin, const float* t_rand, int32_t B, int32_t n, float* z_out,
                             void* stream) {
  if (!near || !far || !lin || !z_out) return set_error("emap_coarse_z: NULL pointer");
  if (B <= 0 || n <= 0) return set_error("emap_coarse_z: bad sizes");
  const long long total = (long long)B * n;
  coarse_z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      near, far, near_is_per_ray ? 1 : 0, lin, t_rand, B, n, z_out);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_upsample_step(const float* rays_o, const float* rays_d, const float* z_in,
                                  const float* udf_in, int32_t n, const float* z_add,
                                  const float* udf_add, int32_t ka, float* z_out, float* udf_out,
                                  const float* u, int32_t k, float* z_new, int64_t* inds_out,
                                  float* weights_out, const float* sample_dist, int32_t B, float inv_s,
                                  float beta, float gamma, const float* gamma_dev, int32_t mode,
                                  int32_t alpha_type, int32_t* status, void* stream) {
  if (!z_in || B <= 0 || n <= 0) return set_error("emap_upsample_step: bad arguments");
  if (n + ka > kMaxSamples) return set_error("emap_upsample_step: more than %d samples per ray", kMaxSamples);
  if (k > kMaxNew || ka > kMaxNew) return set_error("emap_upsample_step: more than %d new samples per step", kMaxNew);
  if (ka > 0 && (!z_add || !z_out)) return set_error("emap_upsample_step: merge needs z_add and z_out");
  if (k > 0 && (!rays_o || !rays_d || !udf_in || !u || !z_new || !sample_dist))
    return set_error("emap_upsample_step: sampling needs rays, udf, u, z_new, sample_dist");
  if (k > 0 && ka > 0 && (!udf_add || !udf_out)) return set_error("emap_upsample_step: merge+sample needs udf_add/udf_out");
  if (mode < 0 || mode > 2 || alpha_type < 0 || alpha_type > 1) return set_error("emap_upsample_step: bad mode");
  if (mode == 2 && ka > 0) return set_error("emap_upsample_step: mode 2 (weights given) does not merge");
  UpsampleArgs a;
  a.rays_o = rays_o; a.rays_d = rays_d; a.z_in = z_in; a.udf_in = udf_in; a.n = n;
  a.z_add = z_add; a.udf_add = udf_add; a.ka = ka; a.z_out = z_out; a.udf_out = udf_out;
  a.u = u; a.z_new = z_new; a.inds = (long long*)inds_out; a.w_out = weights_out;
  a.sample_dist = sample_dist; a.gamma_ptr = gamma_dev; a.B = B; a.k = k; a.inv_s = inv_s; a.beta = beta; a.gamma = gamma;
  a.mode = mode; a.alpha_type = alpha_type; a.status = status;
  const size_t smem = (size_t)kRayWarps * 4 * ray_stride(n + ka) * sizeof(float);
  upsample_step_kernel<<<(B + kRayWarps - 1) / kRayWarps, kRayWarps * 32, smem, (cudaStream_t)stream>>>(a);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_render_prep(const float* z, const float* sample_dist, int32_t B, int32_t n,
                                float* dists, float* mid_z, void* stream) {
  if (!z || !sample_dist || !dists || !mid_z || B <= 0 || n <= 0) return set_error("emap_render_prep: bad arguments");
  const long long total = (long long)B * n;
  render_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, sample_dist, B, n, dists, mid_z);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_render_core_fwd(const float* rays_o, const float* rays_d, const float* mid_z,
                                    const float* dists, const float* udf, const float* grad,
                                    const float* scalars, int32_t B, int32_t n, float cos_anneal_ratio,
                                    float flip_saturation, float near_surface, float sparse_scale,
                                    int32_t use_unbias, int32_t use_norm_grad, int32_t alpha_type,
                                    float* weights, float* alpha, float* grad_flip, float* inside_sphere,
                                    float* grad_mag, float* edge, float* depth, float* normals,
                                    double* partials, float* reduced, int32_t* status, void* stream) {
  if (!rays_o || !rays_d || !mid_z || !dists || !udf || !grad || !scalars || !weights || !grad_flip ||
      !inside_sphere || !grad_mag || !edge || !depth || !normals || !partials || !reduced)
    return set_error("emap_render_core_fwd: NULL pointer");
  if (B <= 0 || n <= 0 || n > kMaxSamples) return set_error("emap_render_core_fwd: bad sizes (n <= %d)", kMaxSamples);
  CoreArgs a;
  a.rays_o = rays_o; a.rays_d = rays_d; a.mid_z = mid_z; a.dists = dists; a.udf = udf; a.grad = grad;
  a.scalars = scalars; a.B = B; a.n = n; a.cos_anneal_ratio = cos_anneal_ratio;
  a.flip_saturation = flip_saturation; a.near_surface = near_surface; a.sparse_scale = sparse_scale;
  a.use_unbias = use_unbias; a.use_norm_grad = use_norm_grad; a.alpha_type = alpha_type;
  a.weights = weights; a.alpha = alpha; a.grad_flip = grad_flip; a.inside_sphere = inside_sphere;
  a.grad_mag = grad_mag; a.edge = edge; a.depth = depth; a.normals = normals; a.partials = partials;
  cudaStream_t st = (cudaStream_t)stream;
  render_core_fwd_kernel<<<(B + kRayWarps - 1) / kRayWarps, kRayWarps * 32,
                           (size_t)kRayWarps * 3 * ray_stride(n) * sizeof(float), st>>>(a);
  EMAP_CUDA(cudaGetLastError());
  render_reduce_kernel<<<1, 1024, 0, st>>>(partials, B, reduced, status);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_render_core_bwd(const float* rays_o, const float* rays_d, const float* mid_z,
                                    const float* dists, const float* udf, const float* grad,
                                    const float* scalars, const float* reduced, int32_t B, int32_t n,
                                    float cos_anneal_ratio, float flip_saturation, float near_surface,
                                    float sparse_scale, int32_t use_unbias, int32_t use_norm_grad,
                                    int32_t alpha_type, const float* d_weights, const float* d_edge,
                                    const float* d_depth, const float* d_normals, const float* d_gerr,
                                    const float* d_gerr_ns, const float* d_sparse, float* d_udf,
                                    float* d_grad, double* partials, float* d_scalars, void* stream) {
  if (!rays_o || !rays_d || !mid_z || !dists || !udf || !grad || !scalars || !reduced || !d_udf ||
      !d_grad || !partials || !d_scalars)
    return set_error("emap_render_core_bwd: NULL pointer");
  if (B <= 0 || n <= 0 || n > kMaxSamples) return set_error("emap_render_core_bwd: bad sizes");
  CoreBwdArgs a;
  a.rays_o = rays_o; a.rays_d = rays_d; a.mid_z = mid_z; a.dists = dists; a.udf = udf; a.grad = grad;
  a.scalars = scalars; a.reduced = reduced; a.B = B; a.n = n; a.cos_anneal_ratio = cos_anneal_ratio;
  a.flip_saturation = flip_saturation; a.near_surface = near_surface; a.sparse_scale = sparse_scale;
  a.use_unbias = use_unbias; a.use_norm_grad = use_norm_grad; a.alpha_type = alpha_type;
  a.d_w = d_weights; a.d_edge = d_edge; a.d_depth = d_depth; a.d_normals = d_normals;
  a.d_gerr = d_gerr; a.d_gerr_ns = d_gerr_ns; a.d_sparse = d_sparse;
  a.d_udf = d_udf; a.d_grad = d_grad; a.partials = partials;
  cudaStream_t st = (cudaStream_t)stream;
  render_core_bwd_kernel<<<(B + kBwdWarps - 1) / kBwdWarps, kBwdWarps * 32,
                           (size_t)kBwdWarps * 7 * ray_stride(n) * sizeof(float), st>>>(a);
  EMAP_CUDA(cudaGetLastError());
  scalar_reduce_kernel<<<1, 1024, 0, st>>>(partials, B, 3, d_scalars);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}
