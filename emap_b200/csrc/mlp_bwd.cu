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
// output-layer pull-back (with dW_8 / db_8) and the final stage -- fixed-order sum of the per-CTA partial
// gradients of mlp_dw.cu + weight-norm backward.
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

// Output layer pull-back, with its own weight gradient.  One warp per point (persistent grid of kTopBlocks
// blocks), lane = 8 consecutive columns:
//   a8 = U8[p].w8 + b8, adot8 = U8[P+p].w8 (tangent along S_g Gbar);  udf = f(a8)/scale, f in {abs, square, id}
//   alpha8 = S_u ubar f'(a8)/scale + (S_u/S_g) f''(a8) adot8 ;  alphadot8 = (S_u/S_g) f'(a8)
//   coef[p] = alpha8 ; coef[P+p] = alphadot8          (scales NULL: S_g = S_u = 1)
//   dW_8 += alpha8 U8[p] + alphadot8 U8[P+p] ;  db_8 += alpha8     -> top_partial[block][kTopStride] (fp32, summed
//   over the blocks in a fixed order by emap_bwd_finish)
__global__ void __launch_bounds__(256) dual_top_kernel(const __half* __restrict__ U8, const float* __restrict__ w8,
                                                       const float* __restrict__ b8p, const float* __restrict__ ubar,
                                                       const float* __restrict__ scales, long long P, int udf_type,
                                                       float scale, float* __restrict__ coef,
                                                       float* __restrict__ top_partial) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warps_total = (long long)gridDim.x * 8;
  float wv[8], accw[8], accb = 0.f;
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w8 + lane * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(w8 + lane * 8 + 4));
    wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w; wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) accw[j] = 0.f;
  const float su = scales ? scales[1] : 1.f, sr = scales ? scales[2] : 1.f, b8 = b8p[0];
  // (two points per warp and iteration -- four loads in flight per lane -- was measured: 0.52 vs 0.33 ms, slower)
  for (long long p = (long long)blockIdx.x * 8 + w; p < P; p += warps_total) {
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(U8 + p * 256 + lane * 8));
    const uint4 qt = __ldg(reinterpret_cast<const uint4*>(U8 + (P + p) * 256 + lane * 8));
    const __half2* hv = reinterpret_cast<const __half2*>(&qv);
    const __half2* ht = reinterpret_cast<const __half2*>(&qt);
    float u[8], ud[8], dv = 0.f, dt = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(hv[j]), b = __half22float2(ht[j]);
      u[2 * j] = a.x; u[2 * j + 1] = a.y; ud[2 * j] = b.x; ud[2 * j + 1] = b.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { dv = fmaf(u[j], wv[j], dv); dt = fmaf(ud[j], wv[j], dt); }
    for (int o = 16; o; o >>= 1) { dv += __shfl_xor_sync(0xffffffffu, dv, o); dt += __shfl_xor_sync(0xffffffffu, dt, o); }
    const float a = dv + b8, adot = dt;
    float f1, f2;
    if (udf_type == 0) { f1 = (a > 0.f) ? 1.f : ((a < 0.f) ? -1.f : 0.f); f2 = 0.f; }
    else if (udf_type == 1) { f1 = 2.f * a; f2 = 2.f; }
    else { f1 = 1.f; f2 = 0.f; }
    const float ub = ubar ? ubar[p] : 0.f;
    const float alpha = su * ub * f1 / scale + sr * f2 * adot, alphadot = sr * f1;
    if (lane == 0) { coef[p] = alpha; coef[P + p] = alphadot; }
#pragma unroll
    for (int j = 0; j < 8; ++j) accw[j] = fmaf(alpha, u[j], fmaf(alphadot, ud[j], accw[j]));
    accb += alpha;
  }
  __shared__ float red[8][kTopStride];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[w][lane * 8 + j] = accw[j];
  if (lane == 0) red[w][256] = accb;
  __syncthreads();
  for (int c = threadIdx.x; c < 257; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][c];
    top_partial[(size_t)blockIdx.x * kTopStride + c] = s;
  }
}

// Last stage: add the per-CTA partial weight / bias gradients (mlp_dw.cu, dual_top_kernel) in a fixed order,
// undo the kernels' PE column order, then the weight-norm backward, scattered into the flat gradient
// (parameters() order: bias, g, v per layer) with the loss scale removed:
//   W = g v/||v||:  dg = <dW, v>/||v|| ;  dv = g/||v|| (dW - <dW, v> v/||v||^2)
// One warp per (layer, output row); a row of dW (<= 256 entries) lives in registers between the two passes.
struct FinishArgs {
  const float* flat;          // parameters
  float* flat_grad;           // out
  const float* partial;       // [n_parts][kDwPartialFloats]
  const float* dbp;           // [n_parts][8][256]
  const float* top;           // [kTopBlocks][kTopStride]
  int n_parts;
  int multires;
  int in_dim[kNumLinear], out_dim[kNumLinear];
  const float* scales;        // cotangent scales of the sweep (scales[3] = 1/S_u) or NULL
  int* status;                // optional: bit EMAP_STATUS_NONFINITE_GRAD is set when a gradient is not finite
};
__device__ __forceinline__ int ref_to_pe_col(int k, int multires) {     // inverse of pe_col_to_ref (common.cuh)
  for (int col = 0; col < 64; ++col)
    if (pe_col_to_ref(col, multires) == k) return col;
  return 0;
}
__global__ void __launch_bounds__(256) finish_kernel(const FinishArgs a) {
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
  const float inv = a.scales ? a.scales[3] : 1.f;
  const float mul = ((l == kSkipLayer) ? 0.70710678118654752f : 1.f) * inv;     // the skip layer's input is [h ; PE]/sqrt2
  const int pe = 3 + 6 * a.multires, hid = kHidden - pe;
  float dwv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = lane + 32 * j;
    float s = 0.f;
    if (k < id) {
      if (l == kNumLinear - 1) {
        for (int b = 0; b < kTopBlocks; ++b) s += a.top[(size_t)b * kTopStride + k];
      } else {
        int off, n, col = k;
        if (l == 0) { off = 0; n = 64; col = ref_to_pe_col(k, a.multires); }
        else if (l < kSkipLayer) { off = 16384 + (l - 1) * 65536; n = 256; }
        else if (l == kSkipLayer) {
          if (k < hid) { off = 16384 + 3 * 65536; n = 256; }
          else { off = 16384 + 4 * 65536; n = 64; col = ref_to_pe_col(k - hid, a.multires); }
        } else { off = 32768 + (l - 1) * 65536; n = 256; }
        const float* src = a.partial + off + (size_t)row * n + col;
#pragma unroll 4
        for (int p = 0; p < a.n_parts; ++p) s += src[(size_t)p * kDwPartialFloats];
      }
    }
    dwv[j] = s * mul;
  }
  float dbv = 0.f;
  if (lane == 0) {
    if (l == kNumLinear - 1) { for (int b = 0; b < kTopBlocks; ++b) dbv += a.top[(size_t)b * kTopStride + 256]; }
    else { for (int p = 0; p < a.n_parts; ++p) dbv += a.dbp[((size_t)p * 8 + l) * 256 + row]; }
    dbv *= inv;
  }
  double ss = 0.0, dot = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = lane + 32 * j;
    if (k < id) { const double vv = v[k]; ss += vv * vv; dot += (double)dwv[j] * vv; }
  }
  for (int o = 16; o; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const double nrm = sqrt(ss);
  const float gi = g[row];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = lane + 32 * j;
    if (k < id) {
      const float o = (float)((double)gi / nrm * ((double)dwv[j] - dot * (double)v[k] / ss));
      gv[k] = o;
      bad |= !isfinite(o);
    }
  }
  if (lane == 0) {
    const float og = (float)(dot / nrm);
    gg[row] = og; gb[row] = dbv;
    bad |= !isfinite(og) || !isfinite(dbv);
  }
  if (a.status && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.status, EMAP_STATUS_NONFINITE_GRAD);
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
                            const float* d_udf, const float* scales, int64_t P, float* coef, void* workspace,
                            void* stream) {
  if (check_net(net)) return 1;
  if (!U8_half || !w8 || !b8 || !coef || !workspace || P <= 0) return set_error("emap_bwd_top: bad arguments");
  float* top = (float*)workspace + (size_t)sm_count() * (kDwPartialFloats + 8 * 256);
  dual_top_kernel<<<kTopBlocks, 256, 0, (cudaStream_t)stream>>>((const __half*)U8_half, w8, b8, d_udf, scales, P,
                                                               net->udf_type, net->scale, coef, top);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_bwd_finish(const emap_net_desc* net, const float* flat_params, const void* workspace,
                               int32_t n_parts, const float* scales, float* flat_grad, int32_t* status,
                               void* stream) {
  if (check_net(net)) return 1;
  if (!flat_params || !workspace || !flat_grad || n_parts <= 0 || n_parts > sm_count())
    return set_error("emap_bwd_finish: bad arguments");
  FinishArgs a;
  a.flat = flat_params; a.flat_grad = flat_grad; a.scales = scales; a.status = status;
  a.partial = (const float*)workspace;
  a.dbp = a.partial + (size_t)sm_count() * kDwPartialFloats;
  a.top = a.dbp + (size_t)sm_count() * 8 * 256;
  a.n_parts = n_parts; a.multires = net->multires;
  net_dims(net->multires, a.in_dim, a.out_dim);
  int rows = 0;
  for (int l = 0; l < kNumLinear; ++l) rows += a.out_dim[l];
  finish_kernel<<<nblk((long long)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(a);
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
