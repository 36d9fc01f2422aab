// K1b: backward of the pair (udf, d udf/dx) w.r.t. the MLP parameters -- the element-wise stages.
//
// The reference obtains these gradients by autograd through two forwards plus
// autograd.grad(create_graph=True) (src/models/udf_model.py:121-135, runner_udf.py:166-167).
// Here the cotangents (ubar = dL/dudf, Gbar = dL/dgrad) are pulled back in closed form.  With
//     phi(theta) = ubar * udf(x) + Gbar . grad_x udf(x) = ubar * f(a8)/scale + f'(a8) * adot8,
// adot = directional derivative of the pre-activation along Gbar, the parameter gradient is the
// reverse sweep of a DUAL forward (value h_l and one tangent hdot_l = sigma_l * adot_l per point):
//     alpha_l    = eta_{l+1} * sigma_l + etadot_{l+1} * adot_l * softplus''(a_l)
//     alphadot_l = etadot_{l+1} * sigma_l
//     dW_l = sum_p alpha_l u_l^T + alphadot_l udot_l^T,   db_l = sum_p alpha_l,
//     [eta_l ; etadot_l] = [alpha_l ; alphadot_l] W_l
// The dual forward / tangent forward (mlp_tc.cu) and the reverse sweep (mlp_rev.cu) are tcgen05 kernels; this
// file holds the small stages around them: the cotangent scales (loss scaling of the fp16 arithmetic), the
// output-layer pull-back, the bias-gradient sums and the weight-norm backward.
#include <math.h>

#include "common.cuh"
#include "host.h"

namespace emap {

// Cotangent scales (loss scaling of the fp16 backward).  The pull-back is jointly linear in
// (ubar = dL/dudf, Gbar = dL/dgrad), but raw loss cotangents at production batch sizes sit far below the
// fp16 range the stashes and MMA operands of the backward are held in (|Gbar| ~ 1e-8: the eikonal term would
// flush to zero).  Two powers of two, computed on the device from the tensors' maxima (no host sync):
//   S_g : the tangent DIRECTION is S_g Gbar (max |S_g Gbar| in [0.5,1)) -- tangent rows hdot ~ O(J);
//   S_u : everything the reverse sweep accumulates is S_u dL/dtheta; the adjoint seed of the tangent output is
//         S_u/S_g, that of the value output S_u ubar;  max(S_u |ubar|, S_u |Gbar|) in [1/8, 1/4).
// scales[0..3] = {S_g, S_u, S_u/S_g, 1/S_u};  scales[4..5] = bit patterns of max|ubar|, max|Gbar| (scratch).
__global__ void __launch_bounds__(256) cotangent_amax_kernel(const float* __restrict__ ubar,
                                                             const float* __restrict__ gbar, long long P,
                                                             unsigned int* __restrict__ amax_bits) {
  float mu = 0.f, mg = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 3 * P; i += stride) {
    if (gbar) mg = fmaxf(mg, fabsf(gbar[i]));      // fmaxf drops NaN: a NaN cotangent surfaces in the result
    if (ubar && i < P) mu = fmaxf(mu, fabsf(ubar[i]));
  }
  for (int o = 16; o; o >>= 1) {
    mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, o));
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
  }
  if ((threadIdx.x & 31) == 0) {                   // non-negative floats order like their bit patterns
    atomicMax(amax_bits + 0, __float_as_uint(mu));
    atomicMax(amax_bits + 1, __float_as_uint(mg));
  }
}
__host__ __device__ inline float pow2_scale(float amax, int target_exp) {
  // power of two S with S * amax in [2^(target_exp-1), 2^target_exp); 1 for 0 / non-finite maxima
  if (!(amax > 0.f) || !(amax < 3.0e38f)) return 1.f;
  int e;
  frexpf(amax, &e);                                // amax = m 2^e, m in [0.5, 1)
  int k = target_exp - e;
  if (k > 100) k = 100;
  if (k < -100) k = -100;
  return ldexpf(1.f, k);
}
__global__ void cotangent_scales_kernel(float* __restrict__ scales) {
  const unsigned int* bits = reinterpret_cast<const unsigned int*>(scales + 4);
  const float au = __uint_as_float(bits[0]), ag = __uint_as_float(bits[1]);
  const float sg = pow2_scale(ag, 0);
  const float su = pow2_scale(fmaxf(au, ag), -2);
  scales[0] = sg; scales[1] = su; scales[2] = su / sg; scales[3] = 1.f / su;
}

// Output layer pull-back.  One warp per point:
//   a8 = U8[p].w8 + b8, adot8 = U8[P+p].w8 (tangent along S_g Gbar);  udf = f(a8)/scale, f in {abs, square, id}
//   alpha8 = S_u ubar f'(a8)/scale + (S_u/S_g) f''(a8) adot8 ;  alphadot8 = (S_u/S_g) f'(a8)
//   coef[p] = alpha8 ; coef[P+p] = alphadot8          (scales NULL: S_g = S_u = 1)
__global__ void dual_top_kernel(const __half* __restrict__ U8, const float* __restrict__ w8,
                                const float* __restrict__ b8p, const float* __restrict__ ubar,
                                const float* __restrict__ scales, long long P, int udf_type, float scale,
                                float* __restrict__ coef) {
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= P) return;
  float dv = 0.f, dt = 0.f;
  for (int c = lane; c < 256; c += 32) {
    const float w = w8[c];
    dv += __half2float(U8[p * 256 + c]) * w;
    dt += __half2float(U8[(P + p) * 256 + c]) * w;
  }
  for (int o = 16; o; o >>= 1) { dv += __shfl_xor_sync(0xffffffffu, dv, o); dt += __shfl_xor_sync(0xffffffffu, dt, o); }
  const float a = dv + b8p[0], adot = dt;
  float f1, f2;
  if (udf_type == 0) { f1 = (a > 0.f) ? 1.f : ((a < 0.f) ? -1.f : 0.f); f2 = 0.f; }
  else if (udf_type == 1) { f1 = 2.f * a; f2 = 2.f; }
  else { f1 = 1.f; f2 = 0.f; }
  const float su = scales ? scales[1] : 1.f, sr = scales ? scales[2] : 1.f;
  const float ub = ubar ? ubar[p] : 0.f;
  if (lane == 0) { coef[p] = su * ub * f1 / scale + sr * f2 * adot; coef[P + p] = sr * f1; }
}

// Weight-norm backward + scatter into the flat gradient (parameters() order: bias, g, v per layer):
//   W = g v/||v||:  dg = <dW, v>/||v|| ;  dv = g/||v|| (dW - <dW, v> v/||v||^2)
// One warp per (layer,row).  dW rows are [out, ldw] fp32 with `col_mul` folded (skip layer 1/sqrt2).
struct WnBwdArgs {
  const float* flat;          // parameters
  float* flat_grad;           // out
  const float* dW[kNumLinear];
  const float* db[kNumLinear];
  int ldw[kNumLinear];
  float mul[kNumLinear];
  int in_dim[kNumLinear], out_dim[kNumLinear];
  const float* scales;        // cotangent scales of the sweep (scales[3] = 1/S_u) or NULL
  int* status;                // optional: bit EMAP_STATUS_NONFINITE_GRAD is set when a gradient is not finite
};
__global__ void wn_bwd_kernel(const WnBwdArgs a) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  int l = 0, row = warp;
  size_t poff = 0;
  for (; l < kNumLinear; ++l) {
    if (row < a.out_dim[l]) break;
    row -= a.out_dim[l];
    poff += (size_t)a.out_dim[l] * (2 + a.in_dim[l]);
  }
  if (l >= kNumLinear) return;
  const int od = a.out_dim[l], id = a.in_dim[l];
  const float* g = a.flat + poff + od;
  const float* v = g + od + (size_t)row * id;
  float* gb = a.flat_grad + poff;
  float* gg = gb + od;
  float* gv = gg + od + (size_t)row * id;
  const float* dW = a.dW[l] + (size_t)row * a.ldw[l];
  const float inv = a.scales ? a.scales[3] : 1.f;
  const float mul = a.mul[l] * inv;
  double ss = 0.0, dot = 0.0;
  for (int k = lane; k < id; k += 32) { const double vv = v[k]; ss += vv * vv; dot += (double)(dW[k] * mul) * vv; }
  for (int o = 16; o; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const double nrm = sqrt(ss);
  const float gi = g[row];
  bool bad = false;
  for (int k = lane; k < id; k += 32) {
    const float o = (float)((double)gi / nrm * ((double)(dW[k] * mul) - dot * (double)v[k] / ss));
    gv[k] = o;
    bad |= !isfinite(o);
  }
  if (lane == 0) {
    const float og = (float)(dot / nrm), ob = a.db[l][row] * inv;
    gg[row] = og; gb[row] = ob;
    bad |= !isfinite(og) || !isfinite(ob);
  }
  if (a.status && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.status, EMAP_STATUS_NONFINITE_GRAD);
}

// Bias gradients db_l[c] = sum_{p < P} A_l[p, c] over the VALUE rows of the eight [2P,256] fp16 stashes the
// reverse sweep wrote (replaces a library reduction that ran at half the HBM rate).  HBM-bound: 4.3 GB at
// P = 1 M.  Pass 1: block b of layer l sums a slab of rows -- a row is 32 lanes x 16 B, a block covers 8 rows
// per step, fp32 accumulation -- into partial[l][b][256]; pass 2 adds the partials in a fixed order:
// deterministic, no atomics.
constexpr int kDbBlocks = 296;          // 2 x 148 slabs per layer
__global__ void __launch_bounds__(256) db_partial_kernel(const __half* __restrict__ st_a, long long P,
                                                         float* __restrict__ partial) {
  const int l = blockIdx.y, b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long rows_per = (P + kDbBlocks - 1) / kDbBlocks;
  const long long r0 = (long long)b * rows_per, r1 = min(P, r0 + rows_per);
  const __half* base = st_a + (size_t)l * 2 * (size_t)P * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (long long r = r0 + w; r < r1; r += 8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)r * 256));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
    }
  }
  __shared__ float red[8][256];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = acc[i];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
  partial[((size_t)l * kDbBlocks + b) * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) db_final_kernel(const float* __restrict__ partial, float* __restrict__ db) {
  const int l = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < kDbBlocks; ++b) s += partial[((size_t)l * kDbBlocks + b) * 256 + threadIdx.x];
  db[l * 256 + threadIdx.x] = s;
}

}  // namespace emap

using namespace emap;

static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" int emap_bwd_cotangent_scales(const float* d_udf, const float* d_grad, int64_t P, float* scales8,
                                         void* stream) {
  if (!scales8 || P <= 0) return set_error("emap_bwd_cotangent_scales: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EMAP_CUDA(cudaMemsetAsync(scales8 + 4, 0, 2 * sizeof(float), st));
  int grid = 2 * sm_count();
  const long long need = (3 * P + 255) / 256;
  if (need < grid) grid = (int)need;
  cotangent_amax_kernel<<<grid, 256, 0, st>>>(d_udf, d_grad, P, reinterpret_cast<unsigned int*>(scales8 + 4));
  EMAP_CUDA(cudaGetLastError());
  cotangent_scales_kernel<<<1, 1, 0, st>>>(scales8);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_bwd_top(const emap_net_desc* net, const void* U8_half, const float* w8, const float* b8,
                            const float* d_udf, const float* scales, int64_t P, float* coef, void* stream) {
  if (check_net(net)) return 1;
  if (!U8_half || !w8 || !b8 || !coef || P <= 0) return set_error("emap_bwd_top: bad arguments");
  dual_top_kernel<<<nblk(P * 32, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)U8_half, w8, b8, d_udf, scales,
                                                                       P, net->udf_type, net->scale, coef);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_bwd_bias_sums(const void* st_a_half, int64_t P, float* partial, float* db, void* stream) {
  if (!st_a_half || !partial || !db || P <= 0) return set_error("emap_bwd_bias_sums: bad arguments");
  db_partial_kernel<<<dim3(kDbBlocks, 8), 256, 0, (cudaStream_t)stream>>>((const __half*)st_a_half, P, partial);
  EMAP_CUDA(cudaGetLastError());
  db_final_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(partial, db);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_bwd_weight_norm(const emap_net_desc* net, const float* flat_params,
                                    const float* const* dW, const int32_t* ldw, const float* mul,
                                    const float* const* db, const float* scales, float* flat_grad,
                                    int32_t* status, void* stream) {
  if (check_net(net)) return 1;
  if (!flat_params || !dW || !ldw || !mul || !db || !flat_grad) return set_error("emap_bwd_weight_norm: NULL pointer");
  WnBwdArgs a;
  a.flat = flat_params; a.flat_grad = flat_grad; a.scales = scales; a.status = status;
  net_dims(net->multires, a.in_dim, a.out_dim);
  int rows = 0;
  for (int l = 0; l < kNumLinear; ++l) {
    if (!dW[l] || !db[l]) return set_error("emap_bwd_weight_norm: NULL layer pointer");
    a.dW[l] = dW[l]; a.db[l] = db[l]; a.ldw[l] = ldw[l]; a.mul[l] = mul[l];
    rows += a.out_dim[l];
  }
  wn_bwd_kernel<<<nblk((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(a);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

// byte offsets inside the packed buffer: out[0] = bias100, out[1..9] = W_eff layer 0..8 (fp32 [out,in])
extern "C" int emap_packed_offsets(const emap_net_desc* net, uint32_t* out10) {
  if (check_net(net)) return 1;
  if (!out10) return set_error("emap_packed_offsets: NULL pointer");
  PackedHeader h; std::vector<RingItem> t1, t3;
  build_layout(*net, h, t1, t3);
  out10[0] = h.bias100_off;
  for (int l = 0; l < kNumLinear; ++l) out10[1 + l] = h.weff_layer_off[l];
  return 0;
}
