// K1b stage 3: the weight-gradient contractions of the backward on the tensor cores, with the bias sums.
//
//     dW_l = A_l^T U_l          l = 0..7      ([256 x 2P] . [2P x 256], layer 0 and the skip term: [2P x 64])
//     db_l = sum_{p < P} A_l[p, :]            (value rows only)
//
// A_l = [alpha_l ; alphadot_l] is the stash the reverse sweep wrote (mlp_rev.cu), U_l = [h_l ; hdot_l] the stash
// of the forward passes (mlp_rg.cu value rows, mlp_tc.cu MODE 3 tangent rows), both fp16 row-major [2P, 256]
// (U_0: [2P, 64], kernel PE column order).  Replaces the `mm` half of autograd's addmm backward
// (src/models/udf_model.py:102 under loss.backward(), runner_udf.py:167) -- nine library GEMMs, their split-K
// reductions and a separate column-sum pass in round 1.
//
// Both operands are "transposed" for the tensor core: the contraction runs over ROWS (points).  A row-major
// [64 rows x 64 columns] fp16 tile landed by the TMA engine with the 128-byte swizzle is exactly the canonical
// MN-major SWIZZLE_128B operand layout of tcgen05.mma (64 MN elements = one 128-byte line per K index, 8 K
// indices = one 1024-byte atom): no transposition anywhere, only the major-ness bits of the instruction
// descriptor.  One persistent CTA per SM owns a contiguous slab of rows and walks the nine jobs over it:
//   warp 8  producer : per 64-row stage, TMA boxes of A_l (4 x 8 KiB) and U_l (4 or 1 x 8 KiB) into a 3-stage ring
//   warp 9  MMA      : per stage 4 (K=16) x 2 (M halves of 128 output features) tcgen05.mma, N = 256 | 64,
//                      fp32 accumulators = 2 x 256 TMEM columns: the whole [256 x 256] dW_l of the slab
//   warps 0-7 (work) : during the main loop the bias column sums of the stage's value rows straight from the A
//                      tile in shared memory; at the end of a job they drain the accumulators into this CTA's
//                      partial result.  emap_bwd_finish adds the partials in a fixed order (deterministic, no
//                      atomics) inside the weight-norm backward.
// HBM-bound by construction: 17.7 GB of stash per 1 M points (the tensor pipe needs 38 % of that time).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "host.h"

namespace emap {
namespace dw {

constexpr int kRows = 64;                         // rows (the contraction dimension) per ring stage
constexpr int kStages = 3;
constexpr int kBoxBytes = kRows * 128;            // one TMA box: [64 rows x 64 columns] fp16, 128B-swizzled
constexpr int kTileBytes = 4 * kBoxBytes;         // [64 x 256]
constexpr int kStageBytes = 2 * kTileBytes;       // A tile | U tile
constexpr int kWorkWarps = 8;
constexpr int kProducerWarp = kWorkWarps, kMmaWarp = kWorkWarps + 1;
constexpr int kThreads = (kWorkWarps + 2) * 32;
constexpr int kJobs = 9;
constexpr int kMaps = 16;                         // 0..7: A_0..A_7;  8..14: st_u[0..6] = U_1..U_7;  15: st_u0 = U_0
constexpr int kPartialFloats = kDwPartialFloats;  // per CTA: 491,520 (host.h: the workspace layout)

struct Smem {
  static constexpr int ring = 0;
  static constexpr int bars = kStages * kStageBytes;
  static constexpr int total = bars + 256 + 1024;
};
static_assert(Smem::total <= 232448, "shared memory plan exceeds 227 KiB");

struct Job { int a_map, u_map, n, off, db_layer; };
// layer 0 | 1 2 3 | 4 (hidden part) | 4 (skip / PE part) | 5 6 7
__constant__ Job c_jobs[kJobs] = {
    {0, 15, 64, 0, 0},
    {1, 8, 256, 16384, 1}, {2, 9, 256, 16384 + 65536, 2}, {3, 10, 256, 16384 + 2 * 65536, 3},
    {4, 11, 256, 16384 + 3 * 65536, 4}, {4, 15, 64, 16384 + 4 * 65536, -1},
    {5, 12, 256, 32768 + 4 * 65536, 5}, {6, 13, 256, 32768 + 5 * 65536, 6}, {7, 14, 256, 32768 + 6 * 65536, 7}};

struct Maps { CUtensorMap m[kMaps]; };

struct Args {
  long long P;
  int stages_total;      // ceil(2P / 64)
  float* partial;        // [grid][kPartialFloats]
  float* db_partial;     // [grid][8][256]
  uint32_t lbo, sbo;     // descriptor strides in bytes (8192 / 1024; overridable for bring-up: emap_set_option)
};

// MN-major SWIZZLE_128B shared-memory operand: 64 MN elements (128 B) contiguous per K index; LBO = byte distance
// between 64-element MN blocks (our 8 KiB boxes), SBO = byte distance between groups of 8 K indices (1024 B).
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;              // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;              // SWIZZLE_128B
  return d;
}
// kind::f16, fp16 x fp16 -> fp32, A and B both MN-major (bits 15 / 16)
__device__ __forceinline__ uint32_t make_idesc_f16_mn(int M, int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) weight_grad_kernel(const __grid_constant__ Maps maps, const Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + Smem::ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;             // [kStages]
  uint64_t* empty = bars + 4;        // [kStages]: 1 tcgen05.commit + kWorkWarps arrivals
  uint64_t* acc_full = bars + 8;
  uint64_t* acc_empty = bars + 9;    // kWorkWarps arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  if (warp == kProducerWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1 + kWorkWarps); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kWorkWarps);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's slab of 64-row stages
  const long long s0 = (long long)blockIdx.x * args.stages_total / gridDim.x;
  const long long s1 = (long long)(blockIdx.x + 1) * args.stages_total / gridDim.x;

  if (warp == kProducerWarp) {
    uint32_t stage = 0, round = 0;
    for (int j = 0; j < kJobs; ++j) {
      const Job job = c_jobs[j];
      const int nbox_u = job.n >> 6;
      const uint32_t bytes = (uint32_t)kTileBytes + (uint32_t)nbox_u * kBoxBytes;
#pragma unroll 1
      for (long long s = s0; s < s1; ++s) {
        if (round > 0) mbar_wait(&empty[stage], (round - 1) & 1, 700 + (int)stage, j);
        if (elect_one()) {
          uint8_t* dst = ring + stage * kStageBytes;
          mbar_arrive_expect_tx(&full[stage], bytes);
          const int r = (int)(s * kRows);
#pragma unroll
          for (int c = 0; c < 4; ++c) tma_load_2d(dst + c * kBoxBytes, &maps.m[job.a_map], c * 64, r, &full[stage]);
          for (int c = 0; c < nbox_u; ++c)
            tma_load_2d(dst + kTileBytes + c * kBoxBytes, &maps.m[job.u_map], c * 64, r, &full[stage]);
        }
        __syncwarp();
        if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
      }
    }
  } else if (warp == kMmaWarp) {
    const uint32_t ring_addr = smem_u32(ring);
    uint32_t stage = 0, round = 0;
    for (int j = 0; j < kJobs; ++j) {
      const Job job = c_jobs[j];
      const uint32_t idesc = make_idesc_f16_mn(128, job.n);
      if (j > 0) mbar_wait(acc_empty, (uint32_t)(j - 1) & 1, 710, j);      // the previous job's accumulators are drained
      tc_fence_after();
#pragma unroll 1
      for (long long s = s0; s < s1; ++s) {
        mbar_wait(&full[stage], round & 1, 720 + (int)stage, j);
        tc_fence_after();
        const uint32_t a_addr = ring_addr + stage * kStageBytes;
        const uint32_t u_addr = a_addr + kTileBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {                  // K = 16 rows per MMA: 16 x 128 B further down the tile
            const uint64_t bdesc = make_sw128_mnmajor_desc(u_addr + k * 2048, args.lbo, args.sbo);
#pragma unroll
            for (int mh = 0; mh < 2; ++mh) {             // output features [128 mh, 128 mh + 128): boxes 2 mh, 2 mh + 1
              const uint64_t adesc = make_sw128_mnmajor_desc(a_addr + mh * 2 * kBoxBytes + k * 2048, args.lbo, args.sbo);
              umma_f16(tmem_base + (uint32_t)mh * 256u, adesc, bdesc, idesc, (s == s0 && k == 0) ? 0u : 1u);
            }
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ===================================== work warps =====================================
    const int t = threadIdx.x;                       // 0..255: the bias column this thread sums
    const int q = warp & 3, hf = warp >> 2;          // TMEM lane quarter, column half
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* part = args.partial + (size_t)blockIdx.x * kPartialFloats;
    float* dbp = args.db_partial + (size_t)blockIdx.x * 8 * 256;
    uint32_t stage = 0, round = 0;
    for (int j = 0; j < kJobs; ++j) {
      const Job job = c_jobs[j];
      float dbacc = 0.f;
      const uint32_t coff = (uint32_t)(t >> 6) * kBoxBytes + (uint32_t)((t & 7) << 1);
      const int g = (t & 63) >> 3;
#pragma unroll 1
      for (long long s = s0; s < s1; ++s) {
        mbar_wait(&full[stage], round & 1, 730 + (int)stage, j);
        const long long r0 = s * kRows;
        if (job.db_layer >= 0 && r0 < args.P) {
          const uint8_t* col = ring + stage * kStageBytes + coff;
          const int nv = (int)min((long long)kRows, args.P - r0);      // value rows of this stage
#pragma unroll 8
          for (int i = 0; i < nv; ++i)
            dbacc += __half2float(*reinterpret_cast<const __half*>(col + i * 128 + (((g ^ (i & 7)) & 7) << 4)));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
      }
      if (job.db_layer >= 0) dbp[job.db_layer * 256 + t] = dbacc;
      // drain: rows = output features 128 mh + 32 q + lane, this warp's column half
      mbar_wait(acc_full, (uint32_t)j & 1, 740, j);
      tc_fence_after();
      const int ncol = job.n >> 1;                   // columns per half: 128 | 32
#pragma unroll 1
      for (int mh = 0; mh < 2; ++mh) {
        float* dst = part + job.off + (size_t)(mh * 128 + q * 32 + lane) * job.n + hf * ncol;
#pragma unroll 1
        for (int c = 0; c < ncol; c += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(lane_taddr + (uint32_t)(mh * 256 + hf * ncol + c), r);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; k += 4)
            *reinterpret_cast<float4*>(dst + c + k) = make_float4(__uint_as_float(r[k]), __uint_as_float(r[k + 1]),
                                                                  __uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

static uint32_t g_lbo = kBoxBytes, g_sbo = 1024;
int set_desc_strides(int which, int v) { if (which == 0) g_lbo = (uint32_t)v; else g_sbo = (uint32_t)v; return 0; }

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// row-major fp16 [rows, cols] -> boxes of [64 rows x 64 columns], 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* m, const void* base, long long rows, int cols) {
  auto enc = encode_fn();
  if (!enc) return set_error("cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)kRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

}  // namespace dw

// Tensor map over one stash of the backward (st_u or st_a: fp16 [8][2P][256], value rows [0,P) and tangent rows
// [P,2P) of each plane) seen as [16 half-planes][P points][256 columns]: boxes of [box_points x 64 columns] with the
// 128-byte swizzle; points beyond P are zero-filled on load and dropped on store (ragged last tile).  Used by the
// reverse sweep (mlp_rev.cu), which stages both stashes through shared memory with the TMA engine.
int make_stash_map(void* map_out, const void* base, long long P, int box_points) {
  auto enc = dw::encode_fn();
  if (!enc) return set_error("cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {256, (cuuint64_t)P, 16};
  cuuint64_t strides[2] = {512, (cuuint64_t)P * 512};
  cuuint32_t box[3] = {64, (cuuint32_t)box_points, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled (stash map) failed (%d)", (int)r);
  return 0;
}

}  // namespace emap

using namespace emap;

extern "C" size_t emap_bwd_workspace_bytes(void) {
  return ((size_t)sm_count() * (kDwPartialFloats + 8 * 256) + (size_t)kTopBlocks * kTopStride) * sizeof(float);
}

static int run_weight_grads(const emap_net_desc* net, const void* st_a, const void* st_u0, const void* st_u,
                            int64_t P, void* workspace, size_t workspace_bytes, void* stream, int* n_parts) {
  if (check_net(net)) return 1;
  if (!st_a || !st_u0 || !st_u || !workspace || P <= 0) return set_error("emap_bwd_weight_grads: bad arguments");
  if (workspace_bytes < emap_bwd_workspace_bytes()) return set_error("emap_bwd_weight_grads: workspace too small");
  if (2 * P > 0x7fffffffLL - 64) return set_error("emap_bwd_weight_grads: too many rows");
  dw::Maps maps;
  const size_t plane = (size_t)2 * (size_t)P * 256 * sizeof(__half);
  for (int l = 0; l < 8; ++l)
    if (dw::make_map(&maps.m[l], (const uint8_t*)st_a + (size_t)l * plane, 2 * P, 256)) return 1;
  for (int l = 0; l < 7; ++l)
    if (dw::make_map(&maps.m[8 + l], (const uint8_t*)st_u + (size_t)l * plane, 2 * P, 256)) return 1;
  if (dw::make_map(&maps.m[15], st_u0, 2 * P, 64)) return 1;
  dw::Args a;
  a.P = P;
  a.stages_total = (int)((2 * P + dw::kRows - 1) / dw::kRows);
  a.partial = (float*)workspace;
  a.db_partial = a.partial + (size_t)sm_count() * dw::kPartialFloats;
  a.lbo = dw::g_lbo; a.sbo = dw::g_sbo;
  int grid = sm_count();
  if (a.stages_total < grid) grid = a.stages_total;      // every CTA owns at least one stage
  static bool attr_done = false;
  if (!attr_done) {
    EMAP_CUDA(cudaFuncSetAttribute(dw::weight_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dw::Smem::total));
    attr_done = true;
  }
  dw::weight_grad_kernel<<<grid, dw::kThreads, dw::Smem::total, (cudaStream_t)stream>>>(maps, a);
  EMAP_CUDA(cudaGetLastError());
  *n_parts = grid;
  return 0;
}

// Returns the number of per-CTA partials written (> 0; emap_bwd_finish needs it), or -1 with emap_last_error() set.
extern "C" int emap_bwd_weight_grads(const emap_net_desc* net, const void* st_a, const void* st_u0, const void* st_u,
                                     int64_t P, void* workspace, size_t workspace_bytes, void* stream) {
  int n_parts = 0;
  if (run_weight_grads(net, st_a, st_u0, st_u, P, workspace, workspace_bytes, stream, &n_parts)) return -1;
  return n_parts;
}
