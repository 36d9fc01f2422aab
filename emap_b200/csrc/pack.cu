// K0: weight-norm fold + packing of the MLP weights into tensor-core operand images.
//
// Replaces the reference's per-forward `torch._weight_norm` (src/models/udf_model.py:74; run 63x per
// render()) by one fold per optimizer step.  Outputs, all in one device buffer ("packed"):
//   * W_eff (fp32, row-major [out,in]) per layer:  W = g * v / ||v||_row
//   * bias100 = 100*b  (the epilogue evaluates softplus on t = 100*a)
//   * for every (layer, K-chunk of 64, N-half) a [rows x 64] K-major, 128B-swizzled 16-bit image
//     of W * 16 (hi part) and of the residual (lo part) -- exactly the byte layout tcgen05.mma
//     expects in shared memory, so the MLP kernel streams them with 1-D bulk copies.
//   * the two "ring item" tables (1-term and 3-term MMA schedules) describing the stream order.
//   * the W^T images of the backward's reverse sweep (hi only) and of the reverse-mode gradient
//     kernel (hi and lo; mlp_rg.cu).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "host.h"

namespace emap {

void net_dims(int multires, int* in_dim, int* out_dim) {
  const int pe = 3 + 6 * multires;
  for (int l = 0; l < kNumLinear; ++l) { in_dim[l] = kHidden; out_dim[l] = kHidden; }
  in_dim[0] = pe;
  out_dim[kSkipLayer - 1] = kHidden - pe;
  out_dim[kNumLinear - 1] = 1;
}

size_t flat_param_count(int multires) {
  int in_dim[kNumLinear], out_dim[kNumLinear];
  net_dims(multires, in_dim, out_dim);
  size_t n = 0;
  for (int l = 0; l < kNumLinear; ++l) n += (size_t)out_dim[l] * (2 + in_dim[l]);
  return n;
}

static inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

constexpr int kRgBigImages = 7 * 4 * 2;          // layers 7..1 x 4 K chunks x (hi, lo)
constexpr int kRgSmallImages = 4 * 2;            // layer 0
constexpr uint32_t kRgBigBytes = 256 * 128, kRgSmallBytes = 64 * 128;
constexpr uint32_t kRgImageBytes = kRgBigImages * kRgBigBytes + kRgSmallImages * kRgSmallBytes;

// Build header + item tables (host).  Deterministic function of (multires, elem_type).
void build_layout(const emap_net_desc& net, PackedHeader& h, std::vector<RingItem>& t1,
                  std::vector<RingItem>& t3) {
  memset(&h, 0, sizeof(h));
  h.magic = 0x50414D45u;  // "EMAP"
  h.multires = (uint32_t)net.multires;
  h.elem_type = (uint32_t)net.elem_type;
  h.scale = net.scale;
  h.udf_type = (uint32_t)net.udf_type;
  int in_dim[kNumLinear], out_dim[kNumLinear];
  net_dims(net.multires, in_dim, out_dim);
  for (int l = 0; l < kNumLinear; ++l) { h.in_dim[l] = in_dim[l]; h.out_dim[l] = out_dim[l]; }

  uint32_t off = align_up(sizeof(PackedHeader), 256);
  h.items_off[0] = off;  off += kMaxItems * sizeof(RingItem);
  h.items_off[1] = off;  off += kMaxItems * sizeof(RingItem);
  h.bias100_off = off;   off += kNumLinear * kHidden * sizeof(float);
  h.weff_off = off;
  for (int l = 0; l < kNumLinear; ++l) {
    h.weff_layer_off[l] = off;
    off += (uint32_t)out_dim[l] * in_dim[l] * sizeof(float);
  }
  off = align_up(off, 1024);
  h.images_off = off;

  t1.clear(); t3.clear();
  uint32_t img = 0;
  for (int l = 0; l < kNumLinear; ++l) {
    int N = kHidden;
    // layer 3 keeps N=256: its zero-padded weight rows make the unused accumulator columns exact
    // zeros (stale TMEM there could hold NaN patterns that would poison layer 4 through 0*NaN)
    if (l == kNumLinear - 1) N = 16;
    int kcs[5]; int nkc = 0;
    // a_chunk 4 = "PE": it is generated into activation chunk 0 (tile start, and again for the skip
    // term of layer 4 after that layer's chunk-0 items have completed -- hence PE comes last there)
    if (l == 0) { kcs[nkc++] = 4; }
    else { for (int c = 0; c < 4; ++c) kcs[nkc++] = c; if (l == kSkipLayer) kcs[nkc++] = 4; }
    struct Half { int off, rows; } halves[2]; int nh = 0;
    if (N <= 128) { halves[nh++] = {0, N}; }
    else { halves[nh++] = {0, 128}; halves[nh++] = {128, N - 128}; }
    for (int ic = 0; ic < nkc; ++ic) {
      // stream order inside a K chunk: part (hi, lo) outer, N half inner -- the two halves of one part
      // land in adjacent ring stages and are consumed by a single N=256 MMA group
      for (int part = 0; part < 2; ++part) {
        for (int ih = 0; ih < nh; ++ih) {
          RingItem it; memset(&it, 0, sizeof(it));
          it.gmem_off = h.images_off + img;
          const uint32_t bytes = (uint32_t)halves[ih].rows * 128u;
          it.bytes16 = (uint16_t)(bytes / 16);
          it.layer = (uint8_t)l;
          it.a_chunk = (uint8_t)kcs[ic];
          it.n_off8 = (uint8_t)(halves[ih].off / 8);
          it.n_rows8 = (uint8_t)(halves[ih].rows / 8);
          it.part = (uint8_t)part;
          uint8_t fl = 0;
          if (ic == 0 && part == 0) fl |= kItemFirstOfAcc;
          if (ih == 0 && part == 0) fl |= kItemWaitA;
          if (ic == 0 && ih == 0 && part == 0) fl |= kItemFirstOfLayer;
          it.flags = fl;
          img += bytes;
          t3.push_back(it);
          if (part == 0) t1.push_back(it);
          // last item of layer 4 that reads activation chunk 0
          if (l == kSkipLayer && kcs[ic] == 0 && ih == nh - 1) {
            if (part == 1) t3.back().flags |= kItemChunk0Done;
            if (part == 0) t1.back().flags |= kItemChunk0Done;
          }
        }
      }
    }
    t1.back().flags |= kItemLastOfLayer;
    t3.back().flags |= kItemLastOfLayer;
  }
  h.images_bytes = img;
  h.n_items[0] = (uint32_t)t1.size();
  h.n_items[1] = (uint32_t)t3.size();
  // reverse-sweep stream (K1b): W_l^T images for l = 7..1, K chunk = 64 OUTPUT neurons, N half = 128
  // INPUT neurons, hi part only (the backward runs single-MMA fp16).  Table right after the images.
  uint32_t roff = align_up(h.images_off + img, 1024);
  h.reserved[0] = roff;                                   // rev item table offset
  h.reserved[1] = 7 * 4 * 2;                              // number of rev items
  h.reserved[2] = align_up(roff + h.reserved[1] * (uint32_t)sizeof(RingItem), 1024);   // rev images offset
  // reverse-mode gradient stream (mlp_rg.cu): W_l^T images hi AND lo, l = 7..1 as [256 x 64] operands
  // (32 KiB), then l = 0 as [64 x 64] operands (8 KiB), in consumption order (layer, K chunk, part).
  h.reserved[3] = align_up(h.reserved[2] + h.reserved[1] * (uint32_t)kStageBytes, 1024);   // rg images offset
  h.reserved[4] = kRgImageBytes;
  h.total_bytes = align_up(h.reserved[3] + kRgImageBytes, 1024);
}

void build_rev_items(const PackedHeader& h, std::vector<RingItem>& tr) {
  tr.clear();
  uint32_t off = h.reserved[2];
  for (int l = 7; l >= 1; --l)
    for (int kc = 0; kc < 4; ++kc)
      for (int nh = 0; nh < 2; ++nh) {
        RingItem it; memset(&it, 0, sizeof(it));
        it.gmem_off = off; it.bytes16 = (uint16_t)(kStageBytes / 16);
        it.layer = (uint8_t)l; it.a_chunk = (uint8_t)kc; it.n_off8 = (uint8_t)(nh * 16); it.n_rows8 = 16;
        tr.push_back(it);
        off += kStageBytes;
      }
}

// ---- kernels -----------------------------------------------------------------------------------
// One warp per output row of one layer: W_eff[row,:] = g[row] * v[row,:] / ||v[row,:]||.
__global__ void wn_fold_kernel(const float* __restrict__ flat, uint8_t* __restrict__ packed,
                               int multires) {
  const PackedHeader* h = reinterpret_cast<const PackedHeader*>(packed);
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  // locate (layer,row)
  int l = 0, row = warp;
  size_t poff = 0;
  for (; l < kNumLinear; ++l) {
    int od = (int)h->out_dim[l], id = (int)h->in_dim[l];
    if (row < od) break;
    row -= od;
    poff += (size_t)od * (2 + id);
  }
  if (l >= kNumLinear) return;
  const int od = (int)h->out_dim[l], id = (int)h->in_dim[l];
  const float* bias = flat + poff;
  const float* g = bias + od;
  const float* v = g + od + (size_t)row * id;
  double ss = 0.0;
  for (int k = lane; k < id; k += 32) { double t = (double)v[k]; ss += t * t; }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  // torch._weight_norm: w = v * (g / norm) in fp32
  const float norm = (float)sqrt(ss);
  const float ratio = g[row] / norm;
  float* W = reinterpret_cast<float*>(packed + h->weff_layer_off[l]) + (size_t)row * id;
  for (int k = lane; k < id; k += 32) W[k] = v[k] * ratio;
  if (lane == 0) {
    float* b100 = reinterpret_cast<float*>(packed + h->bias100_off) + l * kHidden;
    b100[row] = (l == kNumLinear - 1) ? bias[row] : kSoftplusBeta * bias[row];
  }
  // zero the padded tail of the bias row (layer 3: rows >= out_dim; layer 8: rows >= 1)
  if (row == 0) {
    float* b100 = reinterpret_cast<float*>(packed + h->bias100_off) + l * kHidden;
    for (int k = od + lane; k < kHidden; k += 32) b100[k] = 0.f;
  }
}

template <typename T> __device__ __forceinline__ T to_elem(float x);
template <> __device__ __forceinline__ __half to_elem<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 to_elem<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <typename T> __device__ __forceinline__ float from_elem(T x);
template <> __device__ __forceinline__ float from_elem<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float from_elem<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

// One block per item of the 3-term table (which enumerates every image, hi and lo).
template <typename T>
__global__ void pack_images_kernel(uint8_t* __restrict__ packed) {
  const PackedHeader* h = reinterpret_cast<const PackedHeader*>(packed);
  const RingItem it = reinterpret_cast<const RingItem*>(packed + h->items_off[1])[blockIdx.x];
  const int l = it.layer, kc = it.a_chunk, n_off = it.n_off8 * 8, rows = it.n_rows8 * 8;
  const int od = (int)h->out_dim[l], id = (int)h->in_dim[l];
  const int L = (int)h->multires, pe = 3 + 6 * L;
  const float* W = reinterpret_cast<const float*>(packed + h->weff_layer_off[l]);
  T* img = reinterpret_cast<T*>(packed + it.gmem_off);
  const float inv_sqrt2 = 0.70710678118654752440f;
  for (int e = threadIdx.x; e < rows * 64; e += blockDim.x) {
    const int n = e >> 6, kk = e & 63;
    const int o = n_off + n;
    float val = 0.f;
    if (o < od) {
      int src = -1;
      float mul = 1.f;
      if (l == 0) {
        src = pe_col_to_ref(kk, L);
      } else if (l == kSkipLayer) {
        mul = inv_sqrt2;
        if (kc < 4) { int idx = kc * 64 + kk; src = (idx < kHidden - pe) ? idx : -1; }
        else { int r = pe_col_to_ref(kk, L); src = (r >= 0) ? (kHidden - pe) + r : -1; }
      } else {
        src = kc * 64 + kk;
      }
      if (src >= 0 && src < id) val = W[(size_t)o * id + src] * mul * kWeightScale;
    }
    const T hi = to_elem<T>(val);
    const T out = (it.part == 0) ? hi : to_elem<T>(val - from_elem<T>(hi));
    img[sw128_offset(n, kk) >> 1] = out;
  }
}

// One block per reverse item: image[n][kk] = 16 * W_l[out = kc*64+kk][in = nh*128+n]  (x 1/sqrt2 for
// the skip layer, whose PE inputs -- in >= 256-pe -- get no adjoint and are zeroed).
template <typename T>
__global__ void pack_rev_images_kernel(uint8_t* __restrict__ packed) {
  const PackedHeader* h = reinterpret_cast<const PackedHeader*>(packed);
  const RingItem it = reinterpret_cast<const RingItem*>(packed + h->reserved[0])[blockIdx.x];
  const int l = it.layer, kc = it.a_chunk, n_off = it.n_off8 * 8;
  const int od = (int)h->out_dim[l], id = (int)h->in_dim[l];
  const int pe = 3 + 6 * (int)h->multires;
  const float* W = reinterpret_cast<const float*>(packed + h->weff_layer_off[l]);
  T* img = reinterpret_cast<T*>(packed + it.gmem_off);
  const float mul = (l == kSkipLayer) ? 0.70710678118654752440f : 1.f;
  const int in_valid = (l == kSkipLayer) ? (kHidden - pe) : id;
  for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
    const int n = e >> 6, kk = e & 63;
    const int in_idx = n_off + n, out_idx = kc * 64 + kk;
    float val = 0.f;
    if (out_idx < od && in_idx < in_valid) val = W[(size_t)out_idx * id + in_idx] * mul * kWeightScale;
    img[sw128_offset(n, kk) >> 1] = to_elem<T>(val);
  }
}

// Reverse-mode gradient images (mlp_rg.cu).  Image b of the stream: b < 56: layer l = 7 - b/8, K chunk
// kc = (b%8)/2, part b&1, [256 x 64]; b >= 56: layer 0, kc = (b-56)/2, part (b-56)&1, [64 x 64].
//   value(n, kk) = 16 * W_l[out = 64 kc + kk][in = src(n)]        (x 1/sqrt2 for the skip layer)
// i.e. the B operand of  eta_l[n] = sum_out alpha_l[out] W_l[out][n].  src(n) = n, except
//   layer 4: n < out3 : hidden input n;  n >= out3 : PE entry rg_pe_ref(n - (out3-1)) of the skip input
//   layer 0: PE entry rg_pe_ref(n)   (row 0 and rows past the PE width are zero).
// Host-callable so that the CPU tests can build the very same images (emap_debug_rg_image).
__host__ __device__ inline void rg_image_geom(int b, int& l, int& kc, int& part, int& rows, uint32_t& off) {
  if (b < kRgBigImages) { l = 7 - b / 8; kc = (b % 8) / 2; part = b & 1; rows = 256; off = (uint32_t)b * kRgBigBytes; }
  else {
    const int b2 = b - kRgBigImages;
    l = 0; kc = b2 / 2; part = b2 & 1; rows = 64; off = kRgBigImages * kRgBigBytes + (uint32_t)b2 * kRgSmallBytes;
  }
}
__host__ __device__ inline float rg_image_value(const float* W, int l, int od, int id, int multires, int kc,
                                                int n, int kk) {
  const int pe = 3 + 6 * multires, out3 = kHidden - pe;
  const int out_idx = kc * 64 + kk;
  if (out_idx >= od) return 0.f;
  int src;
  float mul = 1.f;
  if (l == 0) {
    src = rg_pe_ref(n, multires);
  } else if (l == kSkipLayer) {
    mul = 0.70710678118654752440f;
    if (n < out3) src = n;
    else { const int r = rg_pe_ref(n - (out3 - 1), multires); src = (r >= 0) ? out3 + r : -1; }
  } else {
    src = n;
  }
  if (src < 0 || src >= id) return 0.f;
  return W[(size_t)out_idx * id + src] * mul * kWeightScale;
}

template <typename T>
__global__ void pack_rg_images_kernel(uint8_t* __restrict__ packed) {
  const PackedHeader* h = reinterpret_cast<const PackedHeader*>(packed);
  int l, kc, part, rows; uint32_t off;
  rg_image_geom((int)blockIdx.x, l, kc, part, rows, off);
  const int od = (int)h->out_dim[l], id = (int)h->in_dim[l];
  const float* W = reinterpret_cast<const float*>(packed + h->weff_layer_off[l]);
  T* img = reinterpret_cast<T*>(packed + h->reserved[3] + off);
  for (int e = threadIdx.x; e < rows * 64; e += blockDim.x) {
    const int n = e >> 6, kk = e & 63;
    const float val = rg_image_value(W, l, od, id, (int)h->multires, kc, n, kk);
    const T hi = to_elem<T>(val);
    img[sw128_offset(n, kk) >> 1] = (part == 0) ? hi : to_elem<T>(val - from_elem<T>(hi));
  }
}

// ---- host entry points -------------------------------------------------------------------------
int check_net(const emap_net_desc* net) {
  if (!net) return set_error("net desc is NULL");
  if (net->multires < 0 || net->multires > kMaxFreq) return set_error("multires must be in [0,10]");
  if (net->udf_type < 0 || net->udf_type > 2) return set_error("udf_type must be 0(abs),1(square),2(sdf)");
  if (net->elem_type < 0 || net->elem_type > 1) return set_error("elem_type must be 0(fp16) or 1(bf16)");
  if (!(net->scale > 0.f)) return set_error("scale must be > 0");
  return 0;
}

// Immortal per-descriptor host copy of the packed buffer's header and item tables (see emap_wn_fold).
struct LayoutBlob {
  PackedHeader h;
  RingItem *t1, *t3, *tr;
  size_t n1, n3, nr;
};
static const LayoutBlob* layout_blob(const emap_net_desc& net) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int, uint32_t>, LayoutBlob*> cache;
  uint32_t sbits;
  memcpy(&sbits, &net.scale, sizeof(sbits));
  const auto key = std::make_tuple((int)net.multires, (int)net.udf_type, (int)net.elem_type, sbits);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  PackedHeader h; std::vector<RingItem> t1, t3, tr;
  build_layout(net, h, t1, t3);
  build_rev_items(h, tr);
  const size_t items = t1.size() + t3.size() + tr.size();
  const size_t bytes = sizeof(LayoutBlob) + items * sizeof(RingItem) + 64;
  void* mem = nullptr;
  if (cudaHostAlloc(&mem, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    mem = malloc(bytes);            // pageable: still immortal, only not capturable / not asynchronous
    if (!mem) return nullptr;
  }
  LayoutBlob* lb = new (mem) LayoutBlob();
  lb->h = h;
  uintptr_t q = (reinterpret_cast<uintptr_t>(lb + 1) + 15) & ~(uintptr_t)15;
  lb->t1 = reinterpret_cast<RingItem*>(q);
  lb->t3 = lb->t1 + t1.size();
  lb->tr = lb->t3 + t3.size();
  lb->n1 = t1.size(); lb->n3 = t3.size(); lb->nr = tr.size();
  if (lb->n1) memcpy(lb->t1, t1.data(), lb->n1 * sizeof(RingItem));
  if (lb->n3) memcpy(lb->t3, t3.data(), lb->n3 * sizeof(RingItem));
  if (lb->nr) memcpy(lb->tr, tr.data(), lb->nr * sizeof(RingItem));
  cache[key] = lb;
  return lb;
}

}  // namespace emap

using namespace emap;

extern "C" size_t emap_flat_param_count(const emap_net_desc* net) {
  if (check_net(net)) return 0;
  return flat_param_count(net->multires);
}

extern "C" size_t emap_packed_size(const emap_net_desc* net) {
  if (check_net(net)) return 0;
  PackedHeader h; std::vector<RingItem> t1, t3;
  build_layout(*net, h, t1, t3);
  return h.total_bytes;
}

extern "C" int emap_wn_fold(const emap_net_desc* net, const float* flat_params, void* packed,
                            void* stream_) {
  if (check_net(net)) return 1;
  if (!flat_params || !packed) return set_error("emap_wn_fold: NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const LayoutBlob* lb = layout_blob(*net);
  if (!lb) return set_error("emap_wn_fold: cannot allocate the layout tables");
  const PackedHeader& h = lb->h;
  if (lb->n3 > (size_t)kMaxItems) return set_error("internal: item table overflow");
  // header + tables: a pure function of net, kept in an immortal PINNED host blob per descriptor, so that the four
  // small H2D copies are truly asynchronous and stay valid when the call is captured into a CUDA graph (a copy
  // from a stack/vector temporary would be replayed from a dead address)
  uint8_t* p = reinterpret_cast<uint8_t*>(packed);
  EMAP_CUDA(cudaMemcpyAsync(p, &lb->h, sizeof(PackedHeader), cudaMemcpyHostToDevice, stream));
  EMAP_CUDA(cudaMemcpyAsync(p + h.items_off[0], lb->t1, lb->n1 * sizeof(RingItem), cudaMemcpyHostToDevice, stream));
  EMAP_CUDA(cudaMemcpyAsync(p + h.items_off[1], lb->t3, lb->n3 * sizeof(RingItem), cudaMemcpyHostToDevice, stream));
  EMAP_CUDA(cudaMemcpyAsync(p + h.reserved[0], lb->tr, lb->nr * sizeof(RingItem), cudaMemcpyHostToDevice, stream));
  int rows = 0;
  for (int l = 0; l < kNumLinear; ++l) rows += (int)h.out_dim[l];
  const int threads = 256, wpb = threads / 32;
  wn_fold_kernel<<<(rows + wpb - 1) / wpb, threads, 0, stream>>>(flat_params, p, net->multires);
  EMAP_CUDA(cudaGetLastError());
  if (net->elem_type == 0)
    pack_images_kernel<__half><<<(unsigned)lb->n3, 256, 0, stream>>>(p);
  else
    pack_images_kernel<__nv_bfloat16><<<(unsigned)lb->n3, 256, 0, stream>>>(p);
  EMAP_CUDA(cudaGetLastError());
  if (net->elem_type == 0)
    pack_rev_images_kernel<__half><<<(unsigned)lb->nr, 256, 0, stream>>>(p);
  else
    pack_rev_images_kernel<__nv_bfloat16><<<(unsigned)lb->nr, 256, 0, stream>>>(p);
  EMAP_CUDA(cudaGetLastError());
  if (net->elem_type == 0)
    pack_rg_images_kernel<__half><<<kRgBigImages + kRgSmallImages, 256, 0, stream>>>(p);
  else
    pack_rg_images_kernel<__nv_bfloat16><<<kRgBigImages + kRgSmallImages, 256, 0, stream>>>(p);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}

// Test hook (HOST memory, no GPU needed): the un-split fp32 value image b of the reverse-mode gradient
// stream as the pack kernel computes it, from HOST W_eff matrices of the layer (row-major [out,in]).
// out_host: [rows x 64] floats, row-major (n, kk) -- no swizzle.  Returns rows (256 or 64), or -1.
extern "C" int emap_debug_rg_image(const emap_net_desc* net, int b, const float* W_host, float* out_host,
                                   int32_t* layer_kc_part /*[3]*/) {
  if (check_net(net)) return -1;
  if (b < 0 || b >= kRgBigImages + kRgSmallImages || !W_host || !out_host) { set_error("emap_debug_rg_image: bad arguments"); return -1; }
  int l, kc, part, rows; uint32_t off;
  rg_image_geom(b, l, kc, part, rows, off);
  int in_dim[kNumLinear], out_dim[kNumLinear];
  net_dims(net->multires, in_dim, out_dim);
  for (int n = 0; n < rows; ++n)
    for (int kk = 0; kk < 64; ++kk)
      out_host[n * 64 + kk] = rg_image_value(W_host, l, out_dim[l], in_dim[l], net->multires, kc, n, kk);
  if (layer_kc_part) { layer_kc_part[0] = l; layer_kc_part[1] = kc; layer_kc_part[2] = part; }
  return rows;
}
extern "C" int emap_debug_pe_col_to_ref(int col, int multires) { return pe_col_to_ref(col, multires); }
extern "C" int emap_debug_rg_pe_ref(int k, int multires) { return rg_pe_ref(k, multires); }
