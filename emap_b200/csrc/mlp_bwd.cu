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
// Round-1 structure: the per-layer GEMMs ([2P,in]x[in,256], [2P,256]x[256,in], [256,2P]x[2P,in])
// are plain library GEMMs (cuBLAS through torch.mm, fp16 operands, fp32 accumulate/output), and
// everything between them -- dual positional encoding, softplus / sigmoid / softplus'' stages with
// their stashes, the output-layer pull-back, the weight-norm backward -- is the kernels below.
// (DESIGN.md "backward" lists the fused tcgen05 version of this sweep as the next step.)
#include <math.h>

#include "common.cuh"
#include "host.h"

namespace emap {

__device__ __forceinline__ void bwd_load_point(const float* pts, const float* rays_o, const float* rays_d,
                                               const float* z, int n_per_ray, long long idx, float scale,
                                               float (&x)[3]) {
  if (pts) { x[0] = pts[idx * 3]; x[1] = pts[idx * 3 + 1]; x[2] = pts[idx * 3 + 2]; }
  else {
    const long long ray = idx / n_per_ray;
    const float zz = z[idx];
    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(rays_d[ray * 3 + c], zz));
  }
  if (scale != 1.f) for (int c = 0; c < 3; ++c) x[c] = __fmul_rn(x[c], scale);
}

// U0[2P,64] fp16, reference PE column order (embedder.py:26-35), col >= pe zero.
// rows [0,P): gamma(x);   rows [P,2P): J_gamma(x) . Gbar  (Gbar NULL -> zeros)
__global__ void pe_dual_kernel(const float* pts, const float* rays_o, const float* rays_d, const float* z,
                               int n_per_ray, long long P, float scale, int multires,
                               const float* __restrict__ gbar, __half* __restrict__ U0) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float x[3];
  bwd_load_point(pts, rays_o, rays_d, z, n_per_ray, p, scale, x);
  float g[3] = {0.f, 0.f, 0.f};
  if (gbar) { g[0] = gbar[p * 3]; g[1] = gbar[p * 3 + 1]; g[2] = gbar[p * 3 + 2]; }
  __half* rv = U0 + p * 64;
  __half* rt = U0 + (P + p) * 64;
  for (int c = 0; c < 3; ++c) { rv[c] = __float2half_rn(x[c]); rt[c] = __float2half_rn(g[c]); }
  for (int j = 0; j < kMaxFreq; ++j) {
    const float f = (float)(1 << j);
    for (int c = 0; c < 3; ++c) {
      float s = 0.f, co = 0.f, ts = 0.f, tc = 0.f;
      if (j < multires) {
        sincosf(x[c] * f, &s, &co);
        ts = f * co * g[c];
        tc = -f * s * g[c];
      }
      rv[3 + 6 * j + c] = __float2half_rn(s);      rt[3 + 6 * j + c] = __float2half_rn(ts);
      rv[3 + 6 * j + 3 + c] = __float2half_rn(co); rt[3 + 6 * j + 3 + c] = __float2half_rn(tc);
    }
  }
  rv[63] = __float2half_rn(0.f); rt[63] = __float2half_rn(0.f);
}

__device__ __forceinline__ float sp100(float a, float& sig) {
  const float t = kSoftplusBeta * a;
  const float e = __expf(-fabsf(t));
  const float r = 1.0f / (1.0f + e);
  sig = (t >= 0.f) ? r : e * r;
  return (fmaxf(t, 0.f) + log1pf(e)) * 0.01f;
}

// Dual activation stage of layer l (l = 0..7):
//   acc[2P, ld] fp32 = [U_l ; Udot_l] W_l^T   (rows [0,P) value, [P,2P) tangent), n_out valid columns
//   -> Unext[2P,256] fp16 (h ; sigma*adot), sig[P,256] fp16, adot[P,256] fp16
// For the skip layer (l == 3) columns [n_out, 256) of Unext receive the PE (U0 cols [0,pe)), and the
// whole row is NOT scaled: the 1/sqrt(2) lives in the fp16 copy of W_4.
__global__ void dual_act_fwd_kernel(const float* __restrict__ acc, int ld, const float* __restrict__ bias,
                                    long long P, int n_out, const __half* __restrict__ U0, int pe,
                                    __half* __restrict__ Unext, __half* __restrict__ sig_out,
                                    __half* __restrict__ adot_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over P*128 (2 cols each)
  if (idx >= P * 128) return;
  const long long p = idx >> 7;
  const int c = (int)(idx & 127) * 2;
  float hv[2], ht[2], sg[2], ad[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int cc = c + k;
    if (cc < n_out) {
      const float a = acc[p * ld + cc] + bias[cc];
      const float adot = acc[(P + p) * ld + cc];
      float s;
      hv[k] = sp100(a, s);
      sg[k] = s; ad[k] = adot; ht[k] = s * adot;
    } else {
      const int q = cc - n_out;
      hv[k] = (U0 && q < pe) ? __half2float(U0[p * 64 + q]) : 0.f;
      ht[k] = (U0 && q < pe) ? __half2float(U0[(P + p) * 64 + q]) : 0.f;
      sg[k] = 0.f; ad[k] = 0.f;
    }
  }
  *reinterpret_cast<__half2*>(Unext + p * 256 + c) = __floats2half2_rn(hv[0], hv[1]);
  *reinterpret_cast<__half2*>(Unext + (P + p) * 256 + c) = __floats2half2_rn(ht[0], ht[1]);
  *reinterpret_cast<__half2*>(sig_out + p * 256 + c) = __floats2half2_rn(sg[0], sg[1]);
  *reinterpret_cast<__half2*>(adot_out + p * 256 + c) = __floats2half2_rn(ad[0], ad[1]);
}

// Output layer pull-back.  One warp per point:
//   a8 = U8[p].w8 + b8, adot8 = U8[P+p].w8;  udf = f(a8)/scale, f in {abs, square, identity}
//   alpha8 = ubar f'(a8)/scale + f''(a8) adot8 ;  alphadot8 = f'(a8)
//   Eta8[p] = alpha8 * w8 ; Eta8[P+p] = alphadot8 * w8 ; coef[p] = alpha8 ; coef[P+p] = alphadot8
__global__ void dual_top_kernel(const __half* __restrict__ U8, const float* __restrict__ w8,
                                const float* __restrict__ b8p,
                                const float* __restrict__ ubar, long long P, int udf_type, float scale,
                                float* __restrict__ Eta8, float* __restrict__ coef) {
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
  const float ub = ubar ? ubar[p] : 0.f;
  const float alpha = ub * f1 / scale + f2 * adot;
  const float alphadot = f1;
  if (Eta8) {
    for (int c = lane; c < 256; c += 32) {
      const float w = w8[c];
      Eta8[p * 256 + c] = alpha * w;
      Eta8[(P + p) * 256 + c] = alphadot * w;
    }
  }
  if (lane == 0) { coef[p] = alpha; coef[P + p] = alphadot; }
}

// Reverse activation stage of layer l (l = 7..0):
//   eta[2P, ld] fp32 (adjoints of h_{l+1}, hdot_{l+1}; `mul` folds the skip 1/sqrt(2)), n valid cols
//   -> A[2P,256] fp16 = [alpha_l ; alphadot_l]   (columns >= n zero)
__global__ void dual_act_bwd_kernel(const float* __restrict__ eta, int ld, float mul, long long P, int n,
                                    const __half* __restrict__ sig, const __half* __restrict__ adot,
                                    __half* __restrict__ A) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * 128) return;
  const long long p = idx >> 7;
  const int c = (int)(idx & 127) * 2;
  float al[2] = {0.f, 0.f}, ad[2] = {0.f, 0.f};
  const float2 sg = __half22float2(*reinterpret_cast<const __half2*>(sig + p * 256 + c));
  const float2 at = __half22float2(*reinterpret_cast<const __half2*>(adot + p * 256 + c));
  const float sgv[2] = {sg.x, sg.y}, atv[2] = {at.x, at.y};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int cc = c + k;
    if (cc < n) {
      const float e = eta[p * ld + cc] * mul, ed = eta[(P + p) * ld + cc] * mul;
      const float s = sgv[k];
      const float sp2 = kSoftplusBeta * s * (1.0f - s);
      al[k] = e * s + ed * atv[k] * sp2;
      ad[k] = ed * s;
    }
  }
  *reinterpret_cast<__half2*>(A + p * 256 + c) = __floats2half2_rn(al[0], al[1]);
  *reinterpret_cast<__half2*>(A + (P + p) * 256 + c) = __floats2half2_rn(ad[0], ad[1]);
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
  const float mul = a.mul[l];
  double ss = 0.0, dot = 0.0;
  for (int k = lane; k < id; k += 32) { const double vv = v[k]; ss += vv * vv; dot += (double)(dW[k] * mul) * vv; }
  for (int o = 16; o; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const double nrm = sqrt(ss);
  const float gi = g[row];
  for (int k = lane; k < id; k += 32)
    gv[k] = (float)((double)gi / nrm * ((double)(dW[k] * mul) - dot * (double)v[k] / ss));
  if (lane == 0) { gg[row] = (float)(dot / nrm); gb[row] = a.db[l][row]; }
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

extern "C" int emap_bwd_pe_dual(const emap_net_desc* net, const float* pts, const float* rays_o,
                                const float* rays_d, const float* z, int32_t n_per_ray, int64_t P,
                                const float* d_grad, void* U0_half, void* stream) {
  if (check_net(net)) return 1;
  if (!U0_half || P <= 0) return set_error("emap_bwd_pe_dual: bad arguments");
  if (!pts && (!rays_o || !rays_d || !z || n_per_ray <= 0)) return set_error("emap_bwd_pe_dual: no points");
  pe_dual_kernel<<<nblk(P, 128), 128, 0, (cudaStream_t)stream>>>(pts, rays_o, rays_d, z, n_per_ray, P,
                                                                 net->scale, net->multires, d_grad,
                                                                 (__half*)U0_half);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_bwd_act_fwd(const float* acc, int32_t ld, const float* bias, int64_t P, int32_t n_out,
                                const void* U0_half, int32_t pe, void* Unext_half, void* sig_half,
                                void* adot_half, void* stream) {
  if (!acc || !bias || !Unext_half || !sig_half || !adot_half || P <= 0) return set_error("emap_bwd_act_fwd: bad arguments");
  dual_act_fwd_kernel<<<nblk(P * 128, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, ld, bias, P, n_out, (const __half*)U0_half, pe, (__half*)Unext_half, (__half*)sig_half, (__half*)adot_half);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_bwd_top(const emap_net_desc* net, const void* U8_half, const float* w8, const float* b8,
                            const float* d_udf, int64_t P, float* Eta8, float* coef, void* stream) {
  if (check_net(net)) return 1;
  if (!U8_half || !w8 || !b8 || !coef || P <= 0) return set_error("emap_bwd_top: bad arguments");
  dual_top_kernel<<<nblk(P * 32, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)U8_half, w8, b8, d_udf, P,
                                                                       net->udf_type, net->scale, Eta8, coef);
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

extern "C" int emap_bwd_act_bwd(const float* eta, int32_t ld, float mul, int64_t P, int32_t n,
                                const void* sig_half, const void* adot_half, void* A_half, void* stream) {
  if (!eta || !sig_half || !adot_half || !A_half || P <= 0 || n > 256 || ld < n) return set_error("emap_bwd_act_bwd: bad arguments");
  dual_act_bwd_kernel<<<nblk(P * 128, 256), 256, 0, (cudaStream_t)stream>>>(
      eta, ld, mul, P, n, (const __half*)sig_half, (const __half*)adot_half, (__half*)A_half);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_bwd_weight_norm(const emap_net_desc* net, const float* flat_params,
                                    const float* const* dW, const int32_t* ldw, const float* mul,
                                    const float* const* db, float* flat_grad, void* stream) {
  if (check_net(net)) return 1;
  if (!flat_params || !dW || !ldw || !mul || !db || !flat_grad) return set_error("emap_bwd_weight_norm: NULL pointer");
  WnBwdArgs a;
  a.flat = flat_params; a.flat_grad = flat_grad;
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
