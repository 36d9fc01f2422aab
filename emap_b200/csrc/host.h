// Host-side glue shared by the .cu translation units: error reporting for the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>

#include "../../include/emap_b200.h"
#include "common.cuh"

namespace emap {

int set_error(const char* fmt, ...);   // stores message (thread-local), returns 1
int check_net(const emap_net_desc* net);
void net_dims(int multires, int* in_dim, int* out_dim);
size_t flat_param_count(int multires);
void build_layout(const emap_net_desc& net, PackedHeader& h, std::vector<RingItem>& t1,
                  std::vector<RingItem>& t3);
int sm_count();
int make_stash_map(void* map_out_128B, const void* base, long long P, int box_points);   // mlp_dw.cu: TMA view of st_u / st_a
unsigned int* tile_counter(cudaStream_t stream);   // zeroed per-launch counter for dynamic tile scheduling (or NULL)

// Workspace of the backward's weight-gradient stage (emap_bwd_workspace_bytes), in floats:
//   [sm_count][kDwPartialFloats]  per-CTA partial dW of the nine contraction jobs (mlp_dw.cu: c_jobs offsets)
//   [sm_count][8][256]            per-CTA partial db_0..7
//   [kTopBlocks][kTopStride]      per-block partial dW_8[256], db_8 of the output-layer pull-back (mlp_bwd.cu)
constexpr int kDwPartialFloats = 2 * 256 * 64 + 7 * 256 * 256;
constexpr int kTopBlocks = 592;
constexpr int kTopStride = 260;

#define EMAP_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::emap::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
                               cudaGetErrorString(_e));                                          \
  } while (0)

}  // namespace emap
