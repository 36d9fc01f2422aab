// Device helpers shared by the tensor-core MLP kernels (mlp_tc.cu: K1 / K1g / dual forward; mlp_rg.cu:
// forward + reverse-mode gradient): kernel arguments, shared-memory plan, 16-bit element packing into the
// 128-byte-swizzled A tile, softplus/sigmoid, point loading and the positional-encoding input stage.
#pragma once
#include "common.cuh"
#include "host.h"

namespace emap {

constexpr int kEpiWarps = 16;                 // 4 per TMEM lane quarter
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kChunkBytes = 16384;           // one [128 x 64] 16-bit SW128 chunk
constexpr int kScratchFloatsPerWarp = 8 * 20;   // MODE_GRAD: 8 points x 16 columns (+4 pad: conflict-free)

struct MlpArgs {
  const uint8_t* packed;
  const float* pts;
  const float* rays_o;
  const float* rays_d;
  const float* z;
  int n_per_ray;
  long long P;
  float* udf_out;
  float* grad_out;
  float* pe_out;
  // MODE_DUAL (backward recompute): cotangent direction + fp16 stashes for the reverse sweep / dW GEMMs
  const float* gbar;    // [P,3] dL/d(grad udf) or NULL (zero tangent)
  const float* bwd_scales;   // cotangent scales (emap_bwd_cotangent_scales) or NULL: tangent direction = scales[0] * gbar
  __half* st_u0;        // [2P,64]  dual PE in kernel column order (rows [0,P) value, [P,2P) tangent)
  __half* st_u;         // [8][2P,256] inputs of layers 1..8 (h ; hdot).  This is the ONLY per-layer stash:
                        // sigma_l = 1 - exp(-100 h_{l+1}) and adot_l * softplus''(a_l) = 100 hdot_{l+1} (1 - sigma_l)
                        // are recovered from it by the reverse sweep.
  float* dbg_acc;       // optional [9][128][256] dump of tile 0 accumulators (descaled), else NULL
  int num_tiles;
  int iters;
  long long* dbg_clk;   // optional [2][9][8] clock64 stamps of block 0, tile iteration 1 (epilogue warp 0 / MMA role 0)
  int dbg_flags;        // timing experiments only: 1 = no MMA issue, 2 = no weight copies, 4 = no epilogue math
  int dbg_iter;         // tile iteration of block 0 whose timeline is stamped into dbg_clk (emap_set_option("dbg_iter"))
  unsigned int* tile_counter;   // single-CTA launches: tiles grid, grid+1, ... are handed out in arrival order (NULL = static)
  // MODE 3 (tangent forward) with fp16 stashes: CUtensorMap over st_u as [16 half-planes][P][256], boxes of
  // [128 points x 64 columns] (host.h: make_stash_map) -- the stash rows travel through shared memory by TMA
  alignas(64) unsigned char stash_map[128];
};

template <typename T> struct IsFp16 { static constexpr bool value = false; };
template <> struct IsFp16<__half> { static constexpr bool value = true; };

// TMA helpers of the stash traffic (tangent forward in mlp_tc.cu, training forward in mlp_rg.cu; mlp_rev.cu has its own)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* smem_src, const void* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::
                   "l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

constexpr int kRingStageBytes = 2 * kStageBytes;   // one part: [256 x 64] 16-bit SW128 image (N halves adjacent)

template <int NTERMS, int MODE, bool PAIR = false, bool SLOTS = false>
struct SmemPlan {
  // fp32x3: A_hi + A_lo = 128 KiB, ring 3 x 32 KiB; single-MMA modes: A = 64 KiB, ring 4 x 32 KiB.
  // PAIR (cta_group::2): every CTA of a pair stages only ITS N half of each operand -> 16 KiB stages,
  // twice as many of them in the same bytes.
  // (The PE chunk is not kept resident: it is written into activation chunk 0 for layer 0 and
  //  re-generated there for the skip term of layer 4.  The weight schedule is computed, not tabled.
  //  In fp32x3 MODE_GRAD the value->tangent exchange scratch aliases the destination chunk -- see the
  //  epilogue -- so that the third ring stage fits.)
  static constexpr int kStageBytesP = PAIR ? kStageBytes : kRingStageBytes;
  // MODE 3 (tangent forward, single MMA): four 16 KiB slots through which the TMA engine moves the stash rows
  // (value rows in, tangent rows out, see mlp_tc.cu), paid for with one ring stage.
  static constexpr bool kSlots = SLOTS;
  static constexpr int kStages = ((NTERMS == 3 || kSlots) ? 3 : 4) * (PAIR ? 2 : 1);
  static constexpr bool kOwnScratch = (MODE == 1 && NTERMS == 1);
  static constexpr int a_hi = 0;
  static constexpr int a_lo = a_hi + 4 * kChunkBytes;
  static constexpr int ring = a_lo + ((NTERMS == 3) ? 4 * kChunkBytes : 0);
  static constexpr int slots = ring + kStages * kStageBytesP;
  static constexpr int scratch = slots + (kSlots ? 4 * kChunkBytes : 0);
  static constexpr int bars = scratch + (kOwnScratch ? kEpiWarps * kScratchFloatsPerWarp * 4 : 0);
  static constexpr int total = bars + 384 + 1024;   // +1 KiB slack to 1024-align the base
};
static_assert(SmemPlan<3, 1>::total <= 232448 && SmemPlan<3, 0>::total <= 232448 &&
              SmemPlan<1, 1>::total <= 232448 && SmemPlan<3, 2>::total <= 232448 &&
              SmemPlan<3, 1, true>::total <= 232448 && SmemPlan<1, 3, false, true>::total <= 232448,
              "shared memory plan exceeds 227 KiB");

// per-mode schedule constants (MODE 0 forward, 1 forward+grad (4 rows/point), 2 dual forward with
// stashes for the backward (2 rows/point, layers 0..7 only -- the output layer is pulled back by a
// separate small kernel), 3 tangent-only forward of the backward (1 row/point, layers 0..7; the value rows
// of the stash come from the training forward, mlp_rg.cu))
template <int MODE> struct ModeInfo {
  static constexpr int kLayers = (MODE >= 2) ? 8 : 9;          // MMA layers per tile
  static constexpr int kUses0 = (MODE >= 2) ? 4 : 5;           // accumulator-0 uses per tile
  static constexpr int kAPerTile = (MODE >= 2) ? 7 : 8;        // a_ready[0..3] completions per tile
  static constexpr int kPtsPerTile = (MODE == 0 || MODE == 3) ? 128 : ((MODE == 1) ? 32 : 64);
};

template <typename T> struct Elem;
template <> struct Elem<__half> {
  static constexpr int fmt = 0;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int fmt = 1;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
};

// Write 8 consecutive columns (one 16-byte swizzle group) of one row of a chunk: hi (and lo) parts.
template <int NTERMS, typename T>
__device__ __forceinline__ void store_group(uint8_t* chunk_hi, uint8_t* chunk_lo, int row, int gidx,
                                            const float (&v)[8]) {
  const uint32_t off = (uint32_t)row * 128u + (uint32_t)(((gidx ^ (row & 7)) & 7) << 4);
  uint4 hi;
  hi.x = Elem<T>::pack2(v[0], v[1]);
  hi.y = Elem<T>::pack2(v[2], v[3]);
  hi.z = Elem<T>::pack2(v[4], v[5]);
  hi.w = Elem<T>::pack2(v[6], v[7]);
  *reinterpret_cast<uint4*>(chunk_hi + off) = hi;
  if (NTERMS == 3) {
    float2 b0 = Elem<T>::unpack2(hi.x), b1 = Elem<T>::unpack2(hi.y), b2 = Elem<T>::unpack2(hi.z),
           b3 = Elem<T>::unpack2(hi.w);
    uint4 lo;
    lo.x = Elem<T>::pack2(v[0] - b0.x, v[1] - b0.y);
    lo.y = Elem<T>::pack2(v[2] - b1.x, v[3] - b1.y);
    lo.z = Elem<T>::pack2(v[4] - b2.x, v[5] - b2.y);
    lo.w = Elem<T>::pack2(v[6] - b3.x, v[7] - b3.y);
    *reinterpret_cast<uint4*>(chunk_lo + off) = lo;
  }
}

// Write 4 consecutive columns (half of a 16-byte swizzle group): `half4` selects the low/high 8 bytes.
template <int NTERMS, typename T>
__device__ __forceinline__ void store_half_group(uint8_t* chunk_hi, uint8_t* chunk_lo, int row, int gidx,
                                                 int half4, const float (&v)[4]) {
  const uint32_t off = (uint32_t)row * 128u + (uint32_t)(((gidx ^ (row & 7)) & 7) << 4) + (uint32_t)half4 * 8u;
  uint2 hi;
  hi.x = Elem<T>::pack2(v[0], v[1]);
  hi.y = Elem<T>::pack2(v[2], v[3]);
  *reinterpret_cast<uint2*>(chunk_hi + off) = hi;
  if (NTERMS == 3) {
    float2 b0 = Elem<T>::unpack2(hi.x), b1 = Elem<T>::unpack2(hi.y);
    uint2 lo;
    lo.x = Elem<T>::pack2(v[0] - b0.x, v[1] - b0.y);
    lo.y = Elem<T>::pack2(v[2] - b1.x, v[3] - b1.y);
    *reinterpret_cast<uint2*>(chunk_lo + off) = lo;
  }
}

// softplus(a; beta=100) and, optionally, its derivative sigmoid(100 a), from t = 100*a.
// torch: log1p(exp(100 a))/100, identity above threshold 20 (udf_model.py:78) -- identical in fp32:
// for t > 20 the log1p term is < 2.1e-11 and vanishes against a >= 0.2.
template <bool WITH_SIG>
__device__ __forceinline__ float softplus100(float t, float& sig) {
  const float e = ex2_approx(-fabsf(t) * 1.4426950408889634f);  // exp(-|t|) in (0,1]
  const float onepe = 1.f + e;
  const float l2 = lg2_approx(onepe);
  if (WITH_SIG) {
    const float r = rcp_approx(onepe);
    sig = (t >= 0.f) ? r : e * r;
  }
  return fmaf(l2, 0.0069314718055994531f, fmaxf(t, 0.f) * 0.01f);
}

// sin / cos of a positional-encoding argument a = 2^j x_c (embedder.py:26-35).  CUDA's sincosf carries a
// Payne-Hanek slow path for |a| > 105615; inlined 76 times it made the MLP kernels 490 KB of SASS with an
// instruction-cache hit rate of 76 % (ncu, round 2).  The arguments here are bounded (|x| <= 8192 after the
// scene normalisation, j <= 9), so: Cody-Waite reduction by pi/2 with three FMA constants (exact to ~72 bits),
// the quadrant from the rounding magic number, degree-9 / degree-8 minimax polynomials on [-pi/4, pi/4].
// Measured against float64 over x in [-8,8], j = 0..9: max abs error 7.1e-8 (1.2 ulp at 1); <= 1 ulp from
// torch's CPU sin / cos, which the reference uses.
__host__ __device__ __forceinline__ void sincos_pe(float a, float* s, float* c) {
#ifdef __CUDA_ARCH__
  const float t = fmaf(a, 0.636619747f, 12582912.0f);     // 1.5 * 2^23: nearest integer k in the low mantissa bits
  const int q = __float_as_int(t);
  const float k = t - 12582912.0f;
  float r = fmaf(k, -1.57079601e+00f, a);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  const float r2 = r * r;
  float ps = fmaf(2.86567956e-6f, r2, -1.98559923e-4f);
  ps = fmaf(ps, r2, 8.33338592e-3f);
  ps = fmaf(ps, r2, -1.66666672e-1f);
  const float sn = fmaf(ps, r * r2, r);
  float pc = fmaf(2.44677067e-5f, r2, -1.38877297e-3f);
  pc = fmaf(pc, r2, 4.16666567e-2f);
  pc = fmaf(pc, r2, -0.5f);
  const float cs = fmaf(pc, r2, 1.0f);
  const float ss = (q & 1) ? cs : sn, cc = (q & 1) ? sn : cs;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
#else
  sincosf(a, s, c);
#endif
}

// softplus(a; beta = 100) from t = 100 a, also handing out e = exp(-|t|): sigmoid(t) = (t >= 0 ? 1 : e) / (1 + e)
// can be rebuilt from (e, sign t) later -- K1r stashes that instead of sigma, so its forward epilogue needs
// two MUFU per element (ex2, lg2) instead of three and no float->int conversion (XU pipe: 4 -> 2 ops).
__device__ __forceinline__ float softplus100_e(float t, float& e) {
  e = ex2_approx(-fabsf(t) * 1.4426950408889634f);
  const float l2 = lg2_approx(1.f + e);
  return fmaf(l2, 0.0069314718055994531f, fmaxf(t, 0.f) * 0.01f);
}

__device__ __forceinline__ void load_point(const MlpArgs& a, long long idx, float scale, float (&x)[3]) {
  if (idx >= a.P) idx = a.P - 1;
  if (a.pts) {
    x[0] = a.pts[idx * 3 + 0]; x[1] = a.pts[idx * 3 + 1]; x[2] = a.pts[idx * 3 + 2];
  } else {
    const long long ray = idx / a.n_per_ray;
    const float zz = a.z[idx];
    // reference forms pts = o + d*z as two rounded ops (udf_renderer_blending.py:448,812)
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(a.rays_o[ray * 3 + c], __fmul_rn(a.rays_d[ray * 3 + c], zz));
  }
  if (scale != 1.f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __fmul_rn(x[c], scale);
  }
}

// Input stage for one row: the 32 PE columns owned by thread-half HF (kernel column order, see
// common.cuh), written as four 16-byte groups of the PE chunk.
//   MODE 0: row = point:            [x, sin(2^j x_c), cos(2^j x_c)]          (embedder.py:26-35)
//   MODE 1: row = (point, type):    type 0 as above; type c+1 = d/dx_c of it (the tangent seed).
//   MODE 2: row = (point, value|tangent along gb);   MODE 3: row = point, tangent along gb only;
//   MODE 4 (mlp_rg.cu): as MODE 0, and the values also go to the backward's U_0 stash when one is given.
template <int NTERMS, int MODE, typename T, int HF>
__device__ __forceinline__ void pe_stage(const MlpArgs& args, const float (&x)[3], int multires,
                                         int lane, int row, long long pt, long long tile,
                                         bool emit_pe_out, uint8_t* PE_hi, uint8_t* PE_lo,
                                         const float (&gb)[3]) {
  constexpr int npairs = (HF == 0) ? 14 : 16;
  constexpr int qbase = (HF == 0) ? 0 : 14;
  constexpr int vofs = (HF == 0) ? 4 : 0;
  const int ty = lane & 3;
  float vals[32];
  if (MODE == 0 || MODE == 4) {
    if (HF == 0) { vals[0] = x[0]; vals[1] = x[1]; vals[2] = x[2]; vals[3] = 0.f; }
#pragma unroll
    for (int i = 0; i < npairs; ++i) {
      const int qq = qbase + i, j = qq / 3, ax = qq % 3;
      float s = 0.f, c = 0.f;
      if (j < multires) sincos_pe(x[ax] * (float)(1 << j), &s, &c);
      vals[vofs + 2 * i] = s; vals[vofs + 2 * i + 1] = c;
    }
    if (emit_pe_out && args.pe_out && pt < args.P && tile < args.num_tiles) {
      const int pe = 3 + 6 * multires;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int ref = pe_col_to_ref(HF * 32 + k, multires);
        if (ref >= 0) args.pe_out[pt * pe + ref] = vals[k];
      }
    }
  } else if (MODE == 3) {
    // tangent of the PE along gb (one row per point): d/dt gamma(x + t gb)
    if (HF == 0) { vals[0] = gb[0]; vals[1] = gb[1]; vals[2] = gb[2]; vals[3] = 0.f; }
#pragma unroll
    for (int i = 0; i < npairs; ++i) {
      const int qq = qbase + i, j = qq / 3, ax = qq % 3;
      float s = 0.f, c = 0.f;
      const float f = (float)(1 << j);
      if (j < multires) sincos_pe(x[ax] * f, &s, &c);
      vals[vofs + 2 * i] = f * c * gb[ax]; vals[vofs + 2 * i + 1] = -f * s * gb[ax];
    }
  } else if (MODE == 2) {
    // dual rows: lane pair (value, tangent along gb); the two lanes split the sincos work
    const int t2 = lane & 1;
    float ls[8], lc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = 2 * r + t2;
      const int qq = qbase + i, j = qq / 3, ax = qq - 3 * j;
      const float xa = (ax == 0) ? x[0] : (ax == 1 ? x[1] : x[2]);
      ls[r] = 0.f; lc[r] = 0.f;
      if (i < npairs && j < multires) sincos_pe(xa * (float)(1 << j), &ls[r], &lc[r]);
    }
    if (HF == 0) {
      vals[0] = t2 ? gb[0] : x[0];
      vals[1] = t2 ? gb[1] : x[1];
      vals[2] = t2 ? gb[2] : x[2];
      vals[3] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < npairs; ++i) {
      const int src = (lane & ~1) | (i & 1);
      const float S = __shfl_sync(0xffffffffu, ls[i >> 1], src);
      const float C = __shfl_sync(0xffffffffu, lc[i >> 1], src);
      const int qq = qbase + i, j = qq / 3, ax = qq % 3;
      const float f = (float)(1 << j);
      const float ga = (ax == 0) ? gb[0] : (ax == 1 ? gb[1] : gb[2]);
      float vs = S, vc = C;
      if (t2) { vs = (j < multires) ? f * C * ga : 0.f; vc = (j < multires) ? -f * S * ga : 0.f; }
      vals[vofs + 2 * i] = vs; vals[vofs + 2 * i + 1] = vc;
    }
    if (emit_pe_out && args.st_u0 && pt < args.P && tile < args.num_tiles) {
      __half* dst = args.st_u0 + ((t2 ? args.P : 0) + pt) * 64 + HF * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = Elem<__half>::pack2(vals[g * 8 + 0], vals[g * 8 + 1]);
        v.y = Elem<__half>::pack2(vals[g * 8 + 2], vals[g * 8 + 3]);
        v.z = Elem<__half>::pack2(vals[g * 8 + 4], vals[g * 8 + 5]);
        v.w = Elem<__half>::pack2(vals[g * 8 + 6], vals[g * 8 + 7]);
        *reinterpret_cast<uint4*>(dst + g * 8) = v;
      }
    }
  } else {
    // the 4 lanes of a point split the sincos work, then exchange by shuffle
    float ls[4], lc[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      // lane type ty evaluates pair i = 4r + ty: same instruction stream, per-lane argument
      const int i = 4 * r + ty;
      const int qq = qbase + i, j = qq / 3, ax = qq - 3 * j;
      const float xa = (ax == 0) ? x[0] : (ax == 1 ? x[1] : x[2]);
      ls[r] = 0.f; lc[r] = 0.f;
      if (i < npairs && j < multires) sincos_pe(xa * (float)(1 << j), &ls[r], &lc[r]);
    }
    if (HF == 0) {
      vals[0] = (ty == 0) ? x[0] : (ty == 1 ? 1.f : 0.f);
      vals[1] = (ty == 0) ? x[1] : (ty == 2 ? 1.f : 0.f);
      vals[2] = (ty == 0) ? x[2] : (ty == 3 ? 1.f : 0.f);
      vals[3] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < npairs; ++i) {
      const int src = (lane & ~3) | (i & 3);
      const float S = __shfl_sync(0xffffffffu, ls[i >> 2], src);
      const float C = __shfl_sync(0xffffffffu, lc[i >> 2], src);
      const int qq = qbase + i, j = qq / 3, ax = qq % 3;
      const float f = (float)(1 << j);
      float vs, vc;
      if (ty == 0) { vs = S; vc = C; }
      else if (ax == ty - 1 && j < multires) { vs = f * C; vc = -f * S; }
      else { vs = 0.f; vc = 0.f; }
      vals[vofs + 2 * i] = vs; vals[vofs + 2 * i + 1] = vc;
    }
  }
  if ((MODE == 4 || MODE == 3) && emit_pe_out && args.st_u0 && pt < args.P && tile < args.num_tiles) {
    // backward stash U_0 [2P,64], kernel column order: value rows [0,P) from the training forward (MODE 4 =
    // MODE 0 values + this stash; mlp_rg.cu), tangent rows [P,2P) from the tangent-only forward (MODE 3)
    __half* dst = args.st_u0 + ((MODE == 3 ? args.P : 0) + pt) * 64 + HF * 32;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 v;
      v.x = Elem<__half>::pack2(vals[g * 8 + 0], vals[g * 8 + 1]);
      v.y = Elem<__half>::pack2(vals[g * 8 + 2], vals[g * 8 + 3]);
      v.z = Elem<__half>::pack2(vals[g * 8 + 4], vals[g * 8 + 5]);
      v.w = Elem<__half>::pack2(vals[g * 8 + 6], vals[g * 8 + 7]);
      *reinterpret_cast<uint4*>(dst + g * 8) = v;
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float v8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v8[j] = vals[g * 8 + j];
    store_group<NTERMS, T>(PE_hi, PE_lo, row, HF * 4 + g, v8);
  }
}

}  // namespace emap
