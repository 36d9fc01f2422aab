// the two callers either side of the render path.
//
//   emap_null_direction   -- the "line direction" of edge extraction: for every near-surface voxel the
//       reference stacks sampling_N normalised UDF gradients taken around the voxel into a [S,3] matrix
//       and keeps the right-singular vector of its smallest singular value
//       (src/edge_extraction/extract_pointcloud.py:75-89, :176-187: torch.linalg.svd + vh[:, -1, :] +
//       F.normalize).  Here: one thread per voxel accumulates the 3x3 Gram matrix in fp64 and
//       diagonalises it with cyclic Jacobi rotations -- no [M,S,S] U factor, no cuSOLVER batch.
//   emap_rays_from_pixels -- the deterministic half of the ray sampler
//       (src/dataset/dataset.py:268-305 gen_random_rays_patches_at): pixel -> ndc uv, edge-map gather,
//       K^-1 p, normalise, depth scale, R v, camera centre, fused in one pass over the batch.
//
// Both are HBM/latency-bound element-wise kernels (<= 600 B in / 12 B out per voxel; 16 B in / 52 B out
// per ray): coalesced loads, grid sized to the SM count, nothing to reshape into a GEMM.
#include "common.cuh"
#include "host.h"

namespace emap {

__global__ void null_direction_kernel(const float* __restrict__ grad, long long M, int S,
                                      float* __restrict__ out) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M;
       m += (long long)gridDim.x * blockDim.x) {
    const float* g = grad + m * (long long)S * 3;
    double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0;
    for (int s = 0; s < S; ++s) {
      const double x = g[3 * s], y = g[3 * s + 1], z = g[3 * s + 2];
      a00 += x * x; a01 += x * y; a02 += x * z; a11 += y * y; a12 += y * z; a22 += z * z;
    }
    double a[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
    for (int sweep = 0; sweep < 12; ++sweep) {
      const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
      if (off <= 1e-300 || off <= 1e-18 * (fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]))) break;
#pragma unroll
      for (int pq = 0; pq < 3; ++pq) {
        const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
        const double apq = a[p][q];
        if (fabs(apq) <= 1e-300) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = ((theta >= 0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {          // A <- A J   (columns p, q)
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - sn * akq; a[k][q] = sn * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {          // A <- J^T A (rows p, q)
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - sn * aqk; a[q][k] = sn * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {          // V <- V J
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - sn * vkq; v[k][q] = sn * vkp + c * vkq;
        }
      }
    }
    int j = 0;
    if (a[1][1] < a[j][j]) j = 1;
    if (a[2][2] < a[j][j]) j = 2;
    double x = v[0][j], y = v[1][j], z = v[2][j];
    // F.normalize(., dim=1): v / max(||v||, 1e-12)
    const double nrm = fmax(sqrt(x * x + y * y + z * z), 1e-12);
    out[m * 3 + 0] = (float)(x / nrm); out[m * 3 + 1] = (float)(y / nrm); out[m * 3 + 2] = (float)(z / nrm);
  }
}

struct RayGenArgs {
  const long long* px; const long long* py;
  const float* edge_img;     // [H, W] (the [H,W,1] edge map of the chosen image)
  float kinv[9];             // intrinsics_all_inv[img, :3, :3]
  float rot[9];              // pose_all[img, :3, :3]
  float cen[3];              // pose_all[img, :3, 3]
  int B, H, W;
  float* rays_o; float* rays_v; float* edge; float* ndc_uv; float* p_cam; float* depth_scale;
};

__global__ void rays_from_pixels_kernel(const RayGenArgs a) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.B; i += gridDim.x * blockDim.x) {
    const long long x = a.px[i], y = a.py[i];
    // ndc = 2*pix/(dim-1) - 1: integer product, fp32 division, fp32 subtraction (dataset.py:268-270)
    a.ndc_uv[2 * i + 0] = __fsub_rn(__fdiv_rn((float)(2 * x), (float)(a.W - 1)), 1.f);
    a.ndc_uv[2 * i + 1] = __fsub_rn(__fdiv_rn((float)(2 * y), (float)(a.H - 1)), 1.f);
    a.edge[i] = a.edge_img[y * a.W + x];
    const float p0 = (float)x, p1 = (float)y, p2 = 1.f;
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) pc[r] = a.kinv[3 * r] * p0 + a.kinv[3 * r + 1] * p1 + a.kinv[3 * r + 2] * p2;
    const float nrm = sqrtf(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
    const float v0 = pc[0] / nrm, v1 = pc[1] / nrm, v2 = pc[2] / nrm;
    a.depth_scale[i] = v2;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      a.p_cam[3 * i + r] = pc[r];
      a.rays_v[3 * i + r] = a.rot[3 * r] * v0 + a.rot[3 * r + 1] * v1 + a.rot[3 * r + 2] * v2;
      a.rays_o[3 * i + r] = a.cen[r];
    }
  }
}

}  // namespace emap

using namespace emap;

extern "C" int emap_null_direction(const float* grad, int64_t M, int32_t S, float* out, void* stream) {
  if (!grad || !out) return set_error("emap_null_direction: NULL pointer");
  if (M < 0 || S <= 0) return set_error("emap_null_direction: bad sizes");
  if (M == 0) return 0;
  const int threads = 128;
  long long blocks = (M + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  null_direction_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(grad, M, S, out);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emap_rays_from_pixels(const int64_t* pixels_x, const int64_t* pixels_y, const float* edge_img,
                                     int32_t H, int32_t W, const float* intr_inv3x3, const float* pose4x4,
                                     int32_t B, float* rays_o, float* rays_v, float* edge, float* ndc_uv,
                                     float* p_cam, float* depth_scale, void* stream) {
  if (!pixels_x || !pixels_y || !edge_img || !intr_inv3x3 || !pose4x4 || !rays_o || !rays_v || !edge ||
      !ndc_uv || !p_cam || !depth_scale)
    return set_error("emap_rays_from_pixels: NULL pointer");
  if (B < 0 || H < 2 || W < 2) return set_error("emap_rays_from_pixels: bad sizes");
  if (B == 0) return 0;
  RayGenArgs a;
  a.px = (const long long*)pixels_x; a.py = (const long long*)pixels_y; a.edge_img = edge_img;
  for (int i = 0; i < 9; ++i) a.kinv[i] = intr_inv3x3[i];           // HOST pointers: 9 + 16 floats
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a.rot[3 * r + c] = pose4x4[4 * r + c];
    a.cen[r] = pose4x4[4 * r + 3];
  }
  a.B = B; a.H = H; a.W = W;
  a.rays_o = rays_o; a.rays_v = rays_v; a.edge = edge; a.ndc_uv = ndc_uv; a.p_cam = p_cam;
  a.depth_scale = depth_scale;
  const int threads = 256;
  int blocks = (B + threads - 1) / threads;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  rays_from_pixels_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}
