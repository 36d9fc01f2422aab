// RenderingNetwork.forward as a standalone operator.
//
// reference: src/models/udf_model.py:138-209.  The class is defined and configured (confs/*.conf
// rendering_network: 4 x 128 hidden, mode "no_normal", multires_view 4) but never instantiated or called by the
// reference (runner_base.py:9-13 does not import it; render_core uses a constant edge of ones,
// udf_renderer_blending.py:561) -- so this is NOT part of render() and must not be wired into it.  It exists so that
// a configuration which does use it finds the operator: a fused forward (input assembly incl. the view-direction
// encoding, all Linear + ReLU layers, sigmoid) in fp32 on the CUDA cores.  85,888 MAC per point, activations in
// shared memory, weights (pre-folded and transposed by the host shim, 0.35 MB) read through L1/L2.  Forward only.
#include <math.h>

#include "common.cuh"
#include "host.h"

namespace emap {
namespace rn {

constexpr int kPts = 32;          // points per block
constexpr int kMaxDim = 512;      // widest layer supported
constexpr int kMaxLayers = 8;

struct Args {
  const float* wt[kMaxLayers];    // W_l^T: [in_l, out_l] row-major (weight-norm already folded)
  const float* b[kMaxLayers];
  int dims[kMaxLayers + 1];
  int n_layers;
  int mode;                       // 0 "idr", 1 "no_view_dir", 2 "no_normal"   (udf_model.py:184-195)
  int multires_view, d_feature, d_out, squeeze_out;
  const float* points; const float* normals; const float* view_dirs; const float* feats;
  long long P;
  float* out;                     // [P, d_out]
};

__global__ void __launch_bounds__(256) rendering_mlp_kernel(const Args a) {
  extern __shared__ float sm[];
  const int ld = kMaxDim + 1;
  float* xa = sm;
  float* xb = sm + kPts * ld;
  const long long p0 = (long long)blockIdx.x * kPts;
  const int tid = threadIdx.x;
  const int d0 = a.dims[0];
  // ---- input row: [points, (embedded) view_dirs, normals, -normals, features] as the mode says
  const int vd = (a.mode == 1) ? 0 : ((a.multires_view > 0) ? 3 + 6 * a.multires_view : 3);
  for (int idx = tid; idx < kPts * d0; idx += blockDim.x) {
    const int p = idx / d0, k = idx - p * d0;
    long long pt = p0 + p;
    if (pt >= a.P) pt = a.P - 1;
    float v;
    int c = k;
    if (c < 3) v = a.points[pt * 3 + c];
    else if ((c -= 3) < vd) {
      if (a.multires_view > 0) {                      // embedder.py:26-35: [x, sin(2^j x), cos(2^j x), ...]
        if (c < 3) v = a.view_dirs[pt * 3 + c];
        else {
          const int j = (c - 3) / 6, r = (c - 3) % 6, ax = r % 3;
          const float arg = a.view_dirs[pt * 3 + ax] * (float)(1 << j);
          v = (r < 3) ? sinf(arg) : cosf(arg);
        }
      } else v = a.view_dirs[pt * 3 + c];
    } else {
      c -= vd;
      if (a.mode != 2) {                              // idr / no_view_dir carry normals and -normals
        if (c < 3) { v = a.normals[pt * 3 + c]; goto done; }
        if (c < 6) { v = -a.normals[pt * 3 + c - 3]; goto done; }
        c -= 6;
      }
      v = a.feats[pt * a.d_feature + c];
    }
  done:
    xa[p * ld + k] = v;
  }
  __syncthreads();
  float* xin = xa;
  float* xout = xb;
  for (int l = 0; l < a.n_layers; ++l) {
    const int in = a.dims[l], out = a.dims[l + 1];
    const float* __restrict__ wt = a.wt[l];
    const float* __restrict__ bias = a.b[l];
    const bool last = (l == a.n_layers - 1);
    for (int idx = tid; idx < kPts * out; idx += blockDim.x) {
      const int p = idx / out, o = idx - p * out;     // consecutive threads: consecutive outputs (coalesced W^T rows)
      float acc = bias[o];
      const float* xr = xin + p * ld;
      for (int k = 0; k < in; ++k) acc = fmaf(xr[k], __ldg(wt + (size_t)k * out + o), acc);
      if (!last) acc = fmaxf(acc, 0.f);
      xout[p * ld + o] = acc;
    }
    __syncthreads();
    float* t = xin; xin = xout; xout = t;
  }
  for (int idx = tid; idx < kPts * a.d_out; idx += blockDim.x) {
    const int p = idx / a.d_out, o = idx - p * a.d_out;
    if (p0 + p < a.P) {
      const float v = xin[p * ld + o];
      a.out[(p0 + p) * a.d_out + o] = a.squeeze_out ? 1.f / (1.f + expf(-v)) : v;
    }
  }
}

}  // namespace rn
}  // namespace emap

using namespace emap;

extern "C" int emap_rendering_network_forward(const float* const* wt, const float* const* bias, const int32_t* dims,
                                              int32_t n_layers, int32_t mode, int32_t multires_view,
                                              int32_t d_feature, int32_t d_out, int32_t squeeze_out,
                                              const float* points, const float* normals, const float* view_dirs,
                                              const float* feats, int64_t P, float* out, void* stream) {
  if (!wt || !bias || !dims || !points || !feats || !out || P <= 0) return set_error("emap_rendering_network_forward: bad arguments");
  if (n_layers < 1 || n_layers > rn::kMaxLayers) return set_error("emap_rendering_network_forward: 1..%d layers", rn::kMaxLayers);
  if (mode < 0 || mode > 2) return set_error("emap_rendering_network_forward: mode must be 0 idr, 1 no_view_dir, 2 no_normal");
  if (mode != 1 && !view_dirs) return set_error("emap_rendering_network_forward: view_dirs required");
  if (mode != 2 && !normals) return set_error("emap_rendering_network_forward: normals required");
  rn::Args a;
  for (int l = 0; l <= n_layers; ++l) {
    if (dims[l] < 1 || dims[l] > rn::kMaxDim) return set_error("emap_rendering_network_forward: layer width 1..%d", rn::kMaxDim);
    a.dims[l] = dims[l];
  }
  const int vd = (mode == 1) ? 0 : ((multires_view > 0) ? 3 + 6 * multires_view : 3);
  if (dims[0] != 3 + vd + (mode != 2 ? 6 : 0) + d_feature) return set_error("emap_rendering_network_forward: dims[0] does not match the mode's input");
  if (d_out < 1 || d_out > dims[n_layers]) return set_error("emap_rendering_network_forward: bad d_out");
  for (int l = 0; l < n_layers; ++l) {
    if (!wt[l] || !bias[l]) return set_error("emap_rendering_network_forward: NULL layer pointer");
    a.wt[l] = wt[l]; a.b[l] = bias[l];
  }
  a.n_layers = n_layers; a.mode = mode; a.multires_view = multires_view; a.d_feature = d_feature; a.d_out = d_out;
  a.squeeze_out = squeeze_out; a.points = points; a.normals = normals; a.view_dirs = view_dirs; a.feats = feats;
  a.P = P; a.out = out;
  const size_t smem = 2 * rn::kPts * (rn::kMaxDim + 1) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    EMAP_CUDA(cudaFuncSetAttribute(rn::rendering_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  rn::rendering_mlp_kernel<<<(unsigned)((P + rn::kPts - 1) / rn::kPts), 256, smem, (cudaStream_t)stream>>>(a);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}
