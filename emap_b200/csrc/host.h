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

#define EMAP_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::emap::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,              \
                               cudaGetErrorString(_e));                                          \
  } while (0)

}  // namespace emap
