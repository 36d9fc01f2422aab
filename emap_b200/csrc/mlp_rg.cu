// K1r: fused PE + 9-layer MLP forward AND reverse-mode d udf / d x in one persistent kernel.
//
// Replaces (reference paths relative to /root/reference), like K1g in mlp_tc.cu:
//   src/models/udf_model.py:90-110    UDFNetwork.forward
//   src/models/udf_model.py:121-135   UDFNetwork.gradient  (autograd.grad of the output w.r.t. x)
//   src/models/udf_renderer_blending.py:448-461  (the two calls in render_core)
//
// K1g carries three tangent rows per point through the network (forward mode: 4 rows x 3 split MMAs =
// 12 F executed per point, F = one MLP forward).  This kernel does what autograd does instead: a value-
// only forward (128 points per tile, 3 F) that keeps sigma_l = softplus'(a_l) of every hidden layer,
// then the adjoint sweep
//     alpha_7 = w_8 . sigma_7,      alpha_{l-1} = (alpha_l W_l) . sigma_{l-1},   l = 7..1,
//     d udf / d x = udf'(a_8) J_gamma(x)^T [ alpha_0 W_0  +  PE columns of alpha_4 W_4 / sqrt2 ]
// on the same tile with the W_l^T operand images (3 F): 6 F executed, half the epilogue conversions.
// sigma (7 stashed layers x 256 x 16-bit fixed point = 3.5 KiB per point) does not fit on chip; every CTA
// owns a fixed 448 KiB scratch slice in global memory that it rewrites tile after tile -- thread-private addresses (each
// thread reads back exactly what it wrote), last written = first read, so the slice lives in the
// 126 MB L2 and DRAM sees little of it.
//
// Skeleton = mlp_tc.cu (same roles, ring, barriers, A-tile format): 16 MMA "steps" per tile --
// steps 0..7 = forward layers 0..7, steps 8..15 = reverse layers 7..0 -- ping-ponging the two
// 256-column TMEM accumulators (buf = step & 1).  The output layer is NOT an MMA step: the sweep is linear
// in its seed, so it starts from the unsigned seed w_8 . sigma_7 straight out of layer 7's epilogue (which
// also forms this thread's share of the dot product a_8 = w_8 . h_8 in fp32), and udf'(a_8) multiplies the
// finished gradient.  Adjoints are held in the A tile scaled by 2^4 so that their fp16 lo parts stay
// normal; the scale rides through the sweep and is removed once at the end.
#include "mlp_dev.cuh"

namespace emap {
long long* dbg_clk_buffer();   // mlp_tc.cu (emap_debug_set_clk_buffer)
int dbg_iter();                 // mlp_tc.cu (emap_set_option("dbg_iter"))
namespace rg {

constexpr int kIoWarp = kEpiWarps + 2;            // 19th warp: TMA stores of the training stash (idle in inference)
constexpr int kThreadsRg = kThreads + 32;
constexpr int kSteps = 16;
constexpr int kLastStep = kSteps - 1;
constexpr uint32_t kUsesPerBuf = 8;               // accumulator uses per tile and buffer (buf = step & 1)
constexpr uint32_t kAPerTile = 15;                // completions of a_ready[0..3] per tile (steps 0..14)
constexpr float kAdjScale = 16.f;                 // power of two (headroom: |alpha| < 4095)
constexpr int kSigmaLayers = 7;                   // sigma_0..sigma_6 (sigma_7 is consumed in registers)
// sigma in [0,1] is stashed in 16 bits of FIXED point (15-bit exp(-|t|) + the sign of t, see enc_e2): what
// matters for the gradient is its absolute error (2^-16), not its relative one -- modelled in
// tests/test_rg_emulation.py: 1.3e-5 on the gradient against 5e-6 with fp32 sigma and 1.7e-4 with fp16 sigma.
// Half the L2 traffic of an fp32 stash, and 148 slices (66 MB) fit the 126 MB L2 together with the weights.
constexpr float kSigmaQ = 32767.f;
constexpr int kSigmaWordsPerCta = kSigmaLayers * 4 * 16 * 256;   // [layer][chunk][warp][32 lanes x 8 words] = 448 KiB

// The positional encoding of a tile is not computed on the tile's critical path: while the reverse steps of
// tile i are MMA-bound, the 16 epilogue warps encode the points of tile i+1 and write the finished operand
// image of its PE chunk ([128 x 64] K-major SW128, hi | lo: 32 KiB) into a per-CTA, double-buffered global
// scratch (L2-resident).  One thread then hands it to the TMA engine: a bulk copy into A-tile chunk 0 as soon as
// the last MMA of tile i has released it, and the same copy again for the skip term of layer 4 (the round-1
// kernel ran 2 x 16 sincosf per point on eight warps at both places: ~28 k of the tile's ~218 k clocks with the
// tensor pipe idle -- profiles/r02_k1r_timeline_before.txt).
constexpr int kPeImageBytes = 2 * kChunkBytes;            // hi | lo
constexpr int kPeBytesPerCta = 2 * kPeImageBytes;         // double buffered

struct Args {
  MlpArgs m;
  uint32_t* scratch;     // [grid][kSigmaWordsPerCta]
  uint8_t* pe_scratch;   // [grid][2][kPeImageBytes]
  unsigned int* tile_counter;   // dynamic scheduling: tiles grid, grid+1, ... are handed out in arrival order
  int flags;             // emap_set_option("rg_flags", bits): kFlagSplitTail
};
constexpr int kFlagSplitTail = 1;   // N-split of each step's last K chunk (as K1g does; A/B switch for bring-up)
constexpr int kFlagL2Persist = 2;   // host side: launch with the sigma scratch as a persisting-L2 access window
constexpr int kFlagDynamic = 8;     // tiles handed out by a global atomic counter instead of the static round robin
constexpr int kFlagRolled = 4;      // host side: select the instantiation with the rolled issuer loop
constexpr int kFlagRegStash = 32;   // A/B switch: training stash stored from registers (round-2 form) instead of by TMA
constexpr int kFlagRolledEpi = 16;  // host side: ... and with the reverse steps' chunk loop rolled as well

template <int NTERMS>
struct Plan {
  static constexpr int kStages = (NTERMS == 3) ? 3 : 4;
  static constexpr int a_hi = 0;
  static constexpr int a_lo = a_hi + 4 * kChunkBytes;
  static constexpr int ring = a_lo + ((NTERMS == 3) ? 4 * kChunkBytes : 0);
  static constexpr int bars = ring + kStages * kRingStageBytes;
  static constexpr int total = bars + 256 + 1024;   // +1 KiB slack to 1024-align the base
};
static_assert(Plan<3>::total <= 232448 && Plan<1>::total <= 232448, "shared memory plan exceeds 227 KiB");

// schedule of step s: K chunks, bytes of one operand part, UMMA N
__device__ __forceinline__ constexpr int step_nkc(int s) { return (s == 0) ? 1 : ((s == kSkipLayer) ? 5 : 4); }
__device__ __forceinline__ constexpr uint32_t step_bytes(int s) {
  return (s == kLastStep) ? 8192u : (uint32_t)kRingStageBytes;
}

// sigma scratch: plain (coherent) 256-bit accesses -- the data is rewritten by this kernel, so the read-only
// path of ldg256 must not be used.  Per element 16 bits (enc_e2 below): round(32767 exp(-|t|)), t = 100 a, ones'-
// complemented when t < 0; sigma = (t >= 0 ? 1 : e) / (1 + e) is rebuilt in the reverse step (one MUFU.RCP there
// instead of MUFU.RCP + F2I in the forward step, which is the longer one).  Both conversions go through the
// 2^23 magic number on the FMA / ALU pipes -- no F2I / I2F (they share the XU pipe with the MUFUs).
__device__ __forceinline__ void ld_words8(const uint32_t* p, uint32_t (&u)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "l"(p)
               : "memory");
}
// Two codes per 32-bit word, ONES' COMPLEMENT in 16 bits: q = round(32767 e) for t >= 0, 0xffff ^ q for t < 0 (bit 15 =
// sign of t either way, and e = 0 keeps its sign).  Encoding a pair costs 2 FFMA (2^23 magic number: the low mantissa
// bits are q) + PRMT (pack the low halves) + PRMT (sign-replicate the top byte of each t: 0xffff / 0) + XOR -- 2.5
// instructions per element on the FMA/ALU pipes (the round-2 form extracted and re-inserted the sign bit by bit: 6).
// (prmt through inline PTX: the __byte_perm intrinsic only documents the 3-bit byte index of each selector nibble;
//  bit 3 = "replicate the sign of the selected byte" is a PTX prmt feature)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t enc_e2(float e0, float t0, float e1, float t1) {
  const uint32_t q0 = __float_as_uint(fmaf(e0, kSigmaQ, 12582912.0f));
  const uint32_t q1 = __float_as_uint(fmaf(e1, kSigmaQ, 12582912.0f));
  const uint32_t pack = prmt(q0, q1, 0x5410u);
  const uint32_t mask = prmt(__float_as_uint(t0), __float_as_uint(t1), 0xFFBBu);
  return pack ^ mask;
}
// sigma / 16 of the two codes of a word (the 1/16 of the operand pre-scale folded into the reciprocal's argument):
// XOR with the sign-replicated top bytes restores q in both halves, PRMT builds 2^23 + q, and
// sigma/16 = (t >= 0 ? 1 : e) / (16 + 16 e) with e = q / 32767.
__device__ __forceinline__ void dec_sigma2(uint32_t w, float& s0, float& s1) {
  const uint32_t u = w ^ prmt(w, 0u, 0xBB99u);
  const float q0 = __uint_as_float(prmt(u, 0x4B000000u, 0x7610u)) - 8388608.0f;
  const float q1 = __uint_as_float(prmt(u, 0x4B000000u, 0x7632u)) - 8388608.0f;
  const float r0 = rcp_approx(fmaf(q0, 16.0f / kSigmaQ, 16.0f));
  const float r1 = rcp_approx(fmaf(q1, 16.0f / kSigmaQ, 16.0f));
  s0 = (w & 0x8000u) ? (q0 * (1.0f / kSigmaQ)) * r0 : r0;
  s1 = (w & 0x80000000u) ? (q1 * (1.0f / kSigmaQ)) * r1 : r1;
}

// J_gamma^T applied to 16 consecutive PE adjoints adj[i] <-> PE slot k = kbase + i (rg_pe_ref order,
// common.cuh; kbase even, slots k < 1 are not PE entries): g += sum_k adj_k d gamma_k / d x.
//   x_c: 1;   sin(f x_c): f cos(f x_c);   cos(f x_c): -f sin(f x_c),   f = 2^j   (embedder.py:26-35)
__host__ __device__ __forceinline__ void pe_adjoint16(const float (&adj)[16], int kbase, const float (&x)[3],
                                                      int multires, float (&g)[3]) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const int k = kbase + i;                    // even; warp-uniform
    if (k < 0) continue;
    if (k == 0) { g[1] += adj[i + 1]; }                       // (not PE, x_1)
    else if (k == 2) { g[2] += adj[i]; g[0] += adj[i + 1]; }  // (x_2, x_0)
    else {
      const int qq = (k - 4) >> 1, j = qq / 3, ax = qq - 3 * j;
      if (j < multires) {
        const float f = (float)(1 << j);
        const float xa = (ax == 0) ? x[0] : (ax == 1 ? x[1] : x[2]);
        float s, c;
        sincos_pe(xa * f, &s, &c);
        const float v = f * (adj[i] * c - adj[i + 1] * s);
        if (ax == 0) g[0] += v; else if (ax == 1) g[1] += v; else g[2] += v;
      }
    }
  }
}

// Encode 16 columns (kernel PE column order, common.cuh) of one point: the thread's two 16-byte groups of the PE
// operand image, written to `img` (hi image; lo image kChunkBytes further) -- a global copy of what the
// round-1 kernel stored into shared memory -- and, in training, to the value rows of the backward's U_0 stash.
template <int NTERMS, typename T>
__device__ __forceinline__ void pe_precompute(const MlpArgs& m, const float (&x)[3], int multires, int row,
                                              int sub, long long pt, bool ok, uint8_t* img) {
  float vals[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = sub * 16 + 2 * i;                   // even kernel column, warp-uniform
    float a = 0.f, b = 0.f;
    if (col == 0) { a = x[0]; b = x[1]; }
    else if (col == 2) { a = x[2]; }
    else {
      const int qq = (col < 32) ? ((col - 4) >> 1) : (14 + ((col - 32) >> 1));
      const int j = qq / 3, ax = qq - 3 * j;
      if (j < multires) {
        const float xa = (ax == 0) ? x[0] : (ax == 1 ? x[1] : x[2]);
        sincos_pe(xa * (float)(1 << j), &a, &b);
      }
    }
    vals[2 * i] = a; vals[2 * i + 1] = b;
  }
  if (m.st_u0 && ok) {
    uint32_t pu[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pu[j] = Elem<__half>::pack2(vals[2 * j], vals[2 * j + 1]);
    stg256_cs(m.st_u0 + (size_t)pt * 64 + sub * 16, pu);
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float v8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v8[j] = vals[g * 8 + j];
    store_group<NTERMS, T>(img, img + kChunkBytes, row, sub * 2 + g, v8);
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ROLL >= 1: the issuer walks the 16 steps in a rolled loop (schedule computed at run time) instead of 16 unrolled
// copies of its body -- 11.8 k of the kernel's 19.5 k SASS instructions were the unrolled issuer, executed by ONE
// thread but competing with the 16 epilogue warps for the SM's instruction cache (stall_no_inst 8.5 % in the ncu
// source view): 7.26 -> 6.35 ms in training mode (profiles/r02_stages_time.txt).  ROLL = 2 also rolls the chunk
// loop of the reverse steps' epilogue.
template <int NTERMS, typename T, int ROLL>
__global__ void __launch_bounds__(kThreadsRg, 1) mlp_rgrad_kernel(const __grid_constant__ Args args) {
  using P = Plan<NTERMS>;
  constexpr int kStages = P::kStages;
  constexpr int kParts = (NTERMS == 3) ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const MlpArgs& m = args.m;
  const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(m.packed);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int multires = (int)hdr->multires;
  const float net_scale = hdr->scale;
  const int udf_type = (int)hdr->udf_type;
  const float* bias100 = reinterpret_cast<const float*>(m.packed + hdr->bias100_off);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P::bars);
  uint64_t* full = bars;                  // [kStages]
  uint64_t* empty = bars + 4;             // [kStages]
  uint64_t* a_ready = bars + 8;           // [5]  (index 4 = PE written into chunk 0)
  uint64_t* acc_full = bars + 13;         // [2 buffers][2 N halves]: columns [0,128) / [128,256) complete
  uint64_t* acc_empty = bars + 17;        // [2]
  uint64_t* c0_free = bars + 19;          // layer 4 has consumed chunk 0 -> the PE image may be copied there again
  uint64_t* pe_done = bars + 20;          // all 16 epilogue warps have written their part of a PE image
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  // Tile schedule: the tile of iteration k is published in sched_tile[k & 1] by completion k of sched_ready
  // (one thread of epilogue warp 0, during step 2 of iteration k-1: every role has consumed completion k-1 by
  // then, so the parity wait cannot alias).  A tile index >= num_tiles ends every role's loop.
  uint64_t* sched_ready = bars + 22;
  volatile int* sched_tile = reinterpret_cast<volatile int*>(bars + 23);
  // Training stash (value rows of the backward's U_{l+1}, fp16 = the hi half of the next step's A tile): with fp16
  // operand images the I/O warp writes chunk c of h_{l+1} from the A tile to the stash with ONE TMA store as soon
  // as the 16 epilogue warps have handed it off (st_ready), and the epilogue waits for that store to have read the
  // chunk (st_done) before it overwrites it a step later -- instead of 32-byte register stores per thread and chunk.
  constexpr bool kTmaSt = IsFp16<T>::value;
  const bool tma_st = kTmaSt && m.st_u != nullptr && !(args.flags & kFlagRegStash);
  // (L2 eviction hints -- evict_last on the sigma scratch, evict_first on the stash stores -- were measured: 6.7-6.8
  //  against 6.0-6.6 ms in training mode, slower; so was a persisting-L2 window on the scratch.  Removed.)
  uint64_t* st_ready = bars + 24;         // [4] 7 completions per tile (forward layers 0..6)
  uint64_t* st_done = bars + 28;          // [4] 7 completions per tile

  if (warp == kProducerWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int c = 0; c < 4; ++c) mbar_init(&a_ready[c], kEpiWarps);
    mbar_init(&a_ready[4], 1);            // one arrive.expect_tx by the thread that issues the PE bulk copy
    mbar_init(pe_done, kEpiWarps);
    mbar_init(sched_ready, 1);
    sched_tile[0] = (int)blockIdx.x;
    for (int b = 0; b < 4; ++b) mbar_init(&acc_full[b], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&acc_empty[b], kEpiWarps);
    mbar_init(c0_free, 1);
    for (int c = 0; c < 4; ++c) { mbar_init(&st_ready[c], kEpiWarps); mbar_init(&st_done[c], 1); }
    fence_barrier_init();
    mbar_arrive(sched_ready);             // completion 0: iteration 0 runs tile blockIdx.x
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_block0 = m.dbg_clk ? clock64() : 0;     // debug entry only: per-block elapsed cycles

  if (warp == kProducerWarp) {
    // ===================================== producer =====================================
    // forward images of layers 0..7 (pack.cu: layer -> K chunk -> hi/lo part; they lead the forward stream)
    // then the reverse stream (rg images, same order); one bulk copy per part, uniform control flow, issued
    // by one elected lane.
    const uint8_t* img_f = m.packed + hdr->images_off;
    const uint8_t* img_r = m.packed + hdr->reserved[3];      // reverse image stream (pack.cu: build_layout)
    uint8_t* ring = smem + P::ring;
    uint32_t stage = 0, round = 0;
    for (int iter = 0;; ++iter) {
      mbar_wait(sched_ready, (uint32_t)iter & 1, 560);
      if (sched_tile[iter & 1] >= m.num_tiles) break;
      uint32_t off = 0;
#pragma unroll 1
      for (int s = 0; s < kSteps; ++s) {
        if (s == 8) off = 0;
        const uint8_t* img = (s < 8) ? img_f : img_r;
        const int nparts = step_nkc(s) * 2;
        const uint32_t bytes = step_bytes(s);
#pragma unroll 1
        for (int ip = 0; ip < nparts; ++ip, off += bytes) {
          if (NTERMS == 1 && (ip & 1)) continue;          // single-MMA mode streams the hi images only
          if (round > 0) mbar_wait(&empty[stage], (round - 1) & 1, 100 + (int)stage, s * 16 + ip);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], bytes);
            bulk_g2s(ring + stage * kRingStageBytes, img + off, bytes, &full[stage]);
          }
          __syncwarp();
          if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer ===================================
    const uint32_t a_hi_addr = smem_u32(smem + P::a_hi);
    const uint32_t a_lo_addr = smem_u32(smem + P::a_lo);
    const uint32_t ring_addr = smem_u32(smem + P::ring);
    const uint32_t idesc256 = make_idesc_f16(128, 256, Elem<T>::fmt);
    const uint32_t idesc64 = make_idesc_f16(128, 64, Elem<T>::fmt);
    uint32_t stage = 0, round = 0;
    for (int iter = 0;; ++iter) {
      mbar_wait(sched_ready, (uint32_t)iter & 1, 561);
      if (sched_tile[iter & 1] >= m.num_tiles) break;
#pragma unroll (ROLL >= 1 ? 1 : kSteps)
      for (int s = 0; s < kSteps; ++s) {
        const int buf = s & 1;
        // timeline of block 0's second tile (emap_debug_rgrad + emap_debug_set_clk_buffer): issuer stamps at
        // dbg_clk[64 + 4 s + {0: step start, 1: accumulator free, 2: first K chunk ready, 3: all MMAs issued}]
        const bool stamp = m.dbg_clk && blockIdx.x == 0 && iter == m.dbg_iter && lane == 0;
        if (stamp) m.dbg_clk[64 + 4 * s + 0] = clock64();
        {
          const uint32_t started = (uint32_t)iter * kUsesPerBuf + (uint32_t)(s >> 1);
          if (started > 0) mbar_wait(&acc_empty[buf], (started - 1) & 1, 200 + buf, s);
        }
        if (stamp) m.dbg_clk[64 + 4 * s + 1] = clock64();
        const int nkc = step_nkc(s);
        const uint32_t idesc = (s == kLastStep) ? idesc64 : idesc256;
        const uint32_t d = tmem_base + (uint32_t)buf * 256u;
        // N-split of the last K chunk (optional, mlp_tc.cu): its MMAs into accumulator columns [0,128) are
        // issued and committed first, so the epilogue starts on chunks 0-1 while the tensor pipe finishes
        // columns [128,256).  Not for steps 0 and 4 (their last K chunk is the PE chunk, which lives in
        // activation chunk 0 -- the epilogue of half 0 would overwrite it under the running MMAs of half 1),
        // not for the N=64 last step, not in single-MMA mode.
        const bool split_tail = (NTERMS == 3) && (args.flags & kFlagSplitTail) && s != 0 && s != kSkipLayer &&
                                s != kLastStep;
#pragma unroll
        for (int ic = 0; ic < nkc; ++ic) {
          const int c = (s == 0) ? 4 : ((ic < 4) ? ic : 4);
          {
            const uint32_t uses = (c == 4) ? (uint32_t)iter * 2u + (s == kSkipLayer ? 1u : 0u)
                                           : (uint32_t)iter * kAPerTile + (uint32_t)(s - 1);
            mbar_wait(&a_ready[c], uses & 1, 300 + c, s);
          }
          if (stamp && ic == 0) m.dbg_clk[64 + 4 * s + 2] = clock64();
          tc_fence_after();
          const uint32_t coff = (c == 4) ? 0u : (uint32_t)c * kChunkBytes;   // PE lives in chunk 0
          const uint64_t ahi = make_sw128_kmajor_desc(a_hi_addr + coff);
          const uint64_t alo = make_sw128_kmajor_desc(a_lo_addr + coff);
          if (split_tail && ic == nkc - 1) {
            uint32_t sidx[kParts];
#pragma unroll
            for (int part = 0; part < kParts; ++part) {
              sidx[part] = stage;
              mbar_wait(&full[stage], round & 1, 400 + (int)stage, s * 16 + ic * 2 + part);
              if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
            }
            tc_fence_after();
            const uint32_t idesc128 = make_idesc_f16(128, 128, Elem<T>::fmt);
            if (elect_one()) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t dh = d + (uint32_t)half * 128u;
                const uint64_t b0 = make_sw128_kmajor_desc(ring_addr + sidx[0] * kRingStageBytes + half * kStageBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(dh, ahi + 2 * k, b0 + 2 * k, idesc128, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(dh, alo + 2 * k, b0 + 2 * k, idesc128, 1u);
                const uint64_t b1 = make_sw128_kmajor_desc(ring_addr + sidx[kParts - 1] * kRingStageBytes + half * kStageBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(dh, ahi + 2 * k, b1 + 2 * k, idesc128, 1u);
                umma_commit(&acc_full[buf * 2 + half]);
              }
#pragma unroll
              for (int part = 0; part < kParts; ++part) umma_commit(&empty[sidx[part]]);
            }
            __syncwarp();
            continue;
          }
#pragma unroll
          for (int part = 0; part < kParts; ++part) {
            mbar_wait(&full[stage], round & 1, 400 + (int)stage, s * 16 + ic * 2 + part);
            tc_fence_after();
            const uint64_t bdesc = make_sw128_kmajor_desc(ring_addr + stage * kRingStageBytes);
            if (elect_one()) {
              if (part == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(d, ahi + 2 * k, bdesc + 2 * k, idesc, (ic == 0 && k == 0) ? 0u : 1u);
                if (NTERMS == 3) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma_f16(d, alo + 2 * k, bdesc + 2 * k, idesc, 1u);
                }
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d, ahi + 2 * k, bdesc + 2 * k, idesc, 1u);
              }
              umma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
          }
          if (s == kSkipLayer && ic == 0) { if (elect_one()) umma_commit(c0_free); __syncwarp(); }
        }
        if (!split_tail) {
          if (elect_one()) { umma_commit(&acc_full[buf * 2]); umma_commit(&acc_full[buf * 2 + 1]); }
          __syncwarp();
        }
        if (stamp) m.dbg_clk[64 + 4 * s + 3] = clock64();
      }
    }
  } else if (warp == kIoWarp) {
    // ===================================== training stash: TMA stores, one thread ========
    if (kTmaSt && tma_st && lane == 0) {
      const uint8_t* A_hi_io = smem + P::a_hi;
      for (int iter = 0;; ++iter) {
        mbar_wait(sched_ready, (uint32_t)iter & 1, 565);
        const int tile = sched_tile[iter & 1];
        if (tile >= m.num_tiles) break;
#pragma unroll 1
        for (int l = 0; l < 7; ++l) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            mbar_wait(&st_ready[c], ((uint32_t)iter * 7u + (uint32_t)l) & 1, 630 + c, l);
            tma_store_3d(A_hi_io + c * kChunkBytes, m.stash_map, c * 64, tile * 128, l * 2);   // value rows of plane l
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(&st_done[c]);                           // the chunk may be overwritten
          }
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    // ===================================== epilogue warps ================================
    // warp = 4*sub + q: q = TMEM lane quarter (rows 32q..32q+31 = points), sub = 16-column slice of every
    // 64-column chunk.  All 16 warps convert the same chunk, so chunk c of the next step's A tile is
    // complete after (c+1)/4 of the epilogue and its MMAs start then.
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* A_hi = smem + P::a_hi;
    uint8_t* A_lo = smem + P::a_lo;
    const float k1 = kSoftplusBeta * kInvWeightScale;
    const float b8 = bias100[8 * kHidden];
    const float* w8 = reinterpret_cast<const float*>(m.packed + hdr->weff_layer_off[8]);
    const int out3 = (int)hdr->out_dim[kSkipLayer - 1];
    // this thread's sigma words (l = 0..6): ((l*4 + chunk)*16 + warp)*256 + lane*8
    uint32_t* sg_base = args.scratch + (size_t)blockIdx.x * kSigmaWordsPerCta + (size_t)warp * 256 + lane * 8;
    auto sg_ptr = [&](int l, int chunk) -> uint32_t* { return sg_base + (size_t)(l * 4 + chunk) * 4096; };
    uint8_t* pe_img = args.pe_scratch + (size_t)blockIdx.x * kPeBytesPerCta;
    // hand PE image `b` to the TMA engine: chunk 0 (hi) and its lo twin; completes a_ready[4]
    auto issue_pe_copy = [&](int b) {
      fence_proxy_async_all();                              // the image was written with ordinary global stores
      constexpr uint32_t bytes = (NTERMS == 3) ? (uint32_t)kPeImageBytes : (uint32_t)kChunkBytes;
      mbar_arrive_expect_tx(&a_ready[4], bytes);
      bulk_g2s(A_hi, pe_img + (size_t)b * kPeImageBytes, kChunkBytes, &a_ready[4]);
      if (NTERMS == 3) bulk_g2s(A_lo, pe_img + (size_t)b * kPeImageBytes + kChunkBytes, kChunkBytes, &a_ready[4]);
    };
    // encode the points of tile iteration `it` into image it & 1 (all 16 warps, 16 columns of a row each)
    auto encode_tile = [&](int it, long long t2) {
      const long long p2 = t2 * 128 + row;
      float xn[3];
      load_point(m, p2, net_scale, xn);
      pe_precompute<NTERMS, T>(m, xn, multires, row, sub, p2, (t2 < m.num_tiles) && (p2 < m.P),
                               pe_img + (size_t)(it & 1) * kPeImageBytes);
      __syncwarp();
      if (lane == 0) mbar_arrive(pe_done);
    };
    encode_tile(0, (long long)blockIdx.x);
    if (warp == 0) {
      mbar_wait(pe_done, 0, 550);
      if (lane == 0) issue_pe_copy(0);
      __syncwarp();
    }
    const bool scheduler = (warp == 0 && lane == 0);
    const bool dynamic = (args.flags & kFlagDynamic) != 0;

    for (int iter = 0;; ++iter) {
      mbar_wait(sched_ready, (uint32_t)iter & 1, 562);
      const long long tile = (long long)sched_tile[iter & 1];
      if (tile >= m.num_tiles) break;
      // the tile after this one: fetched now (an L2 atomic in dynamic mode), published during step 2
      int next_tile = 0;
      if (scheduler) {
        const long long nt = dynamic ? (long long)gridDim.x + (long long)atomicAdd(args.tile_counter, 1u)
                                     : tile + (long long)gridDim.x;
        next_tile = (nt < (long long)m.num_tiles) ? (int)nt : m.num_tiles;
      }
      const long long pt = tile * 128 + row;
      const bool ok = (tile < m.num_tiles) && (pt < m.P);
      // (the PE image of this tile is already on its way into chunk 0: nothing to do at tile start)

      // ------------------------------------------------ forward: hidden layers 0..7 (steps 0..7)
      float dot8 = 0.f;                         // this thread's share of a_8 = w_8 . h_8 (its 64 columns)
#pragma unroll 1
      for (int l = 0; l < 8; ++l) {
        const int buf = l & 1;
        const bool top = (l == 7);              // layer 7: h_8 feeds only the output layer; seed the sweep
        const uint32_t acc_par = ((uint32_t)iter * kUsesPerBuf + (uint32_t)(l >> 1)) & 1;
        // epilogue warp 0 stamps at dbg_clk[4 s + {0: waiting, 1: accumulator complete, 2: chunk 0 handed off, 3: done}]
        const bool stamp = m.dbg_clk && blockIdx.x == 0 && iter == m.dbg_iter && warp == 0 && lane == 0;
        if (stamp) m.dbg_clk[4 * l + 0] = clock64();
        mbar_wait(&acc_full[buf * 2], acc_par, 500 + buf, l);
        tc_fence_after();
        if (stamp) m.dbg_clk[4 * l + 1] = clock64();
        if (l == 2 && scheduler) {              // all 16 warps are past step 1 of this tile, i.e. past its schedule wait
          sched_tile[(iter + 1) & 1] = next_tile;
          mbar_arrive(sched_ready);
        }
        const float* bl = bias100 + l * kHidden;
#pragma unroll 1
        for (int chunk = 0; chunk < 4; ++chunk) {
          uint8_t* dst_hi = A_hi + chunk * kChunkBytes;
          uint8_t* dst_lo = A_lo + chunk * kChunkBytes;
          const int col0 = chunk * 64 + sub * 16;
          if (chunk == 2) { mbar_wait(&acc_full[buf * 2 + 1], acc_par, 505 + buf, l); tc_fence_after(); }
          float4 bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(bl + col0) + i);
          // (fetching chunk c+1 from TMEM while chunk c is converted was measured twice -- round 1 on K1g, round 2
          //  here: 6.68 vs 6.05 ms -- and lost both times: the 16 extra live registers cost more than the latency)
          uint32_t r[16];
          tmem_ld_32x32b_x16(lane_taddr + (uint32_t)(buf * 256 + col0), r);
          tmem_wait_ld();
          // the TMA store of this chunk's previous content (h_l, training stash) has read it
          if (tma_st && l >= 1) mbar_wait(&st_done[chunk], ((uint32_t)iter * 7u + (uint32_t)(l - 1)) & 1, 620 + chunk, l);
          if (m.dbg_acc && tile == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              m.dbg_acc[((size_t)l * 128 + row) * 256 + col0 + k] = __uint_as_float(r[k]) * kInvWeightScale;
          }
          uint32_t* sgp = sg_ptr(top ? 0 : l, chunk);
          uint32_t sw[8];                     // (e, sign t) codes of this thread's 16 columns: sigma_l for the sweep
          uint32_t pu[8];                     // h_{l+1} of this thread's 16 columns as fp16 (training stash) ...
          const bool need_pu = m.st_u != nullptr && (top || !tma_st);   // ... where it is not stored by TMA from the A tile
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const float4 bA = bv[2 * g], bB = bv[2 * g + 1];
            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
            float h[8];
            if (!top) {
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const float t0 = fmaf(__uint_as_float(r[g * 8 + j]), k1, bb[j]);
                const float t1 = fmaf(__uint_as_float(r[g * 8 + j + 1]), k1, bb[j + 1]);
                float e0, e1;
                h[j] = softplus100_e(t0, e0);
                h[j + 1] = softplus100_e(t1, e1);
                sw[g * 4 + (j >> 1)] = enc_e2(e0, t0, e1, t1);
              }
              store_group<NTERMS, T>(dst_hi, dst_lo, row, sub * 2 + g, h);
            } else {
              // a_8 += w_8 . h_8;  unsigned seed of the sweep: alpha_7 = w_8 . sigma_7 (x 2^4)
              const float4 wA = __ldg(reinterpret_cast<const float4*>(w8 + col0 + g * 8));
              const float4 wB = __ldg(reinterpret_cast<const float4*>(w8 + col0 + g * 8 + 4));
              const float ww[8] = {wA.x, wA.y, wA.z, wA.w, wB.x, wB.y, wB.z, wB.w};
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float sg;
                h[j] = softplus100<true>(fmaf(__uint_as_float(r[g * 8 + j]), k1, bb[j]), sg);
                dot8 = fmaf(ww[j], h[j], dot8);
                v[j] = kAdjScale * ww[j] * sg;
              }
              store_group<NTERMS, T>(dst_hi, dst_lo, row, sub * 2 + g, v);
            }
            if (need_pu) {
#pragma unroll
              for (int j = 0; j < 4; ++j) pu[g * 4 + j] = Elem<__half>::pack2(h[2 * j], h[2 * j + 1]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&a_ready[chunk]);
            if (tma_st && !top) mbar_arrive(&st_ready[chunk]);
          }
          if (stamp && chunk == 0) m.dbg_clk[4 * l + 2] = clock64();
          if (!top) stg256(sgp, sw);          // after the hand-off: off the MMA's critical path
          // training: value rows [0,P) of the backward's stash U_{l+1} (emap_bwd_tangent_forward adds the
          // tangent rows later) -- after the hand-off, off the MMA's critical path
          // (by TMA from the A tile where that holds the same fp16 values: every layer but the last, fp16 images)
          if (need_pu && ok) stg256_cs(m.st_u + (size_t)l * 2 * (size_t)m.P * 256 + (size_t)pt * 256 + col0, pu);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        if (stamp) m.dbg_clk[4 * l + 3] = clock64();

        if (l == kSkipLayer - 1 && warp == 0) {
          // skip connection: layer 4 = [h4 ; PE]/sqrt2.  Once its MMAs on chunk 0 are done, the PE image is
          // copied there again as the 5th K chunk of layer 4.
          mbar_wait(c0_free, (uint32_t)iter & 1, 520);
          if (tma_st) mbar_wait(&st_done[0], ((uint32_t)iter * 7u + 3u) & 1, 625);    // h_4's chunk 0 has been stored
          if (lane == 0) issue_pe_copy(iter & 1);
          __syncwarp();
        }
      }

      // ------------------------------------------------ reverse: steps 8..14 = layers 7..1
      float gs[3] = {0.f, 0.f, 0.f};            // this thread's share of J_gamma^T (PE adjoint)
      const float inv_adj = kInvWeightScale / kAdjScale;
#pragma unroll 1
      for (int s = 8; s < kLastStep; ++s) {
        const int l = 15 - s;                   // accumulator = alpha_l W_l  -> alpha_{l-1} = acc . sigma_{l-1}
        const int buf = s & 1;
        // sigma of the current chunk; chunk 0 is fetched before the accumulator wait, chunk c+1 as soon as
        // chunk c's values are consumed (its L2 latency then overlaps the conversions and stores of chunk c)
        uint32_t sgc[8];                        // (e, sign t) codes of sigma_{l-1}, this thread's 16 columns
        auto fetch_sigma = [&](int chunk) { ld_words8(sg_ptr(l - 1, chunk), sgc); };
        fetch_sigma(0);
        float x[3] = {0.f, 0.f, 0.f};           // the point itself: only the skip layer's PE adjoint needs it
        if (l == kSkipLayer) load_point(m, pt, net_scale, x);
        const uint32_t acc_par = ((uint32_t)iter * kUsesPerBuf + (uint32_t)(s >> 1)) & 1;
        const bool stamp = m.dbg_clk && blockIdx.x == 0 && iter == m.dbg_iter && warp == 0 && lane == 0;
        if (stamp) m.dbg_clk[4 * s + 0] = clock64();
        mbar_wait(&acc_full[buf * 2], acc_par, 530 + buf, s);
        tc_fence_after();
        if (stamp) m.dbg_clk[4 * s + 1] = clock64();
#pragma unroll (ROLL >= 2 ? 1 : 4)
        for (int chunk = 0; chunk < 4; ++chunk) {
          const int col0 = chunk * 64 + sub * 16;
          if (chunk == 2) { mbar_wait(&acc_full[buf * 2 + 1], acc_par, 535 + buf, s); tc_fence_after(); }
          uint32_t r[16];
          tmem_ld_32x32b_x16(lane_taddr + (uint32_t)(buf * 256 + col0), r);
          tmem_wait_ld();
          if (m.dbg_acc && tile == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              m.dbg_acc[((size_t)s * 128 + row) * 256 + col0 + k] = __uint_as_float(r[k]) * inv_adj;
          }
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float s0, s1;
            dec_sigma2(sgc[j >> 1], s0, s1);                     // sigma / 16 (kInvWeightScale folded in)
            v[j] = __uint_as_float(r[j]) * s0;
            v[j + 1] = __uint_as_float(r[j + 1]) * s1;
          }
          if (chunk < 3) fetch_sigma(chunk + 1);
          if (chunk == 3 && l == kSkipLayer) {
            // columns n >= out3 of alpha_4 W_4 are the adjoint of the skip input's PE part (slot
            // k = n - (out3-1) >= 1): contract them with J_gamma now; they are not inputs of layer 3.
            const int kbase = col0 - (out3 - 1);
            float adj[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) adj[j] = __uint_as_float(r[j]) * inv_adj;
            pe_adjoint16(adj, kbase, x, multires, gs);
#pragma unroll
            for (int j = 0; j < 16; ++j) if (kbase + j >= 1) v[j] = 0.f;
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float v8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v8[j] = v[g * 8 + j];
            store_group<NTERMS, T>(A_hi + chunk * kChunkBytes, A_lo + chunk * kChunkBytes, row, sub * 2 + g, v8);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_ready[chunk]);
          if (stamp && chunk == 0) m.dbg_clk[4 * s + 2] = clock64();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        if (stamp) m.dbg_clk[4 * s + 3] = clock64();
        // the reverse steps are MMA-bound: the gap after the first one takes the encoding of the NEXT tile
        if (s == 8) {
          mbar_wait(sched_ready, (uint32_t)(iter + 1) & 1, 563);      // published in step 2: immediate
          const long long nt = (long long)sched_tile[(iter + 1) & 1];
          if (nt < m.num_tiles) encode_tile(iter + 1, nt);
        }
      }

      // ------------------------------------------------ step 15: alpha_0 W_0 (64 PE slots) -> d udf / d x
      {
        float x[3];
        load_point(m, pt, net_scale, x);
        mbar_wait(&acc_full[2], ((uint32_t)iter * kUsesPerBuf + 7u) & 1, 540);   // buf 1 (both halves commit together)
        tc_fence_after();
        // every MMA of this tile has completed: chunk 0 is free, the next tile's PE image can go in now
        if (warp == 0 && sched_tile[(iter + 1) & 1] < m.num_tiles) {
          mbar_wait(pe_done, (uint32_t)(iter + 1) & 1, 551);
          if (lane == 0) issue_pe_copy((iter + 1) & 1);
          __syncwarp();
        }
        uint32_t r[16];
        tmem_ld_32x32b_x16(lane_taddr + (uint32_t)(256 + sub * 16), r);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[1]);
        float adj[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) adj[j] = __uint_as_float(r[j]) * inv_adj;
        if (m.dbg_acc && tile == 0) {
#pragma unroll
          for (int k = 0; k < 16; ++k) m.dbg_acc[((size_t)kLastStep * 128 + row) * 256 + sub * 16 + k] = adj[k];
        }
        pe_adjoint16(adj, sub * 16, x, multires, gs);
        // sum the four column slices of a point (gradient shares and the a_8 shares): the A tile is dead
        // here (every MMA of the tile has completed), 16 bytes per (point, sub) serve as the exchange --
        // inside the chunk-3 rows of this lane quarter (rows 32q..32q+15), which only its own four warps
        // write later; they meet on a named barrier before and after.
        float4* slots = reinterpret_cast<float4*>(A_hi + 3 * kChunkBytes + (q * 32 + (lane >> 1)) * 128 + (lane & 1) * 64);
        slots[sub] = make_float4(gs[0], gs[1], gs[2], dot8);
        named_bar_sync(1 + q, 128);
        if (sub == 0 && ok) {
          const float4 s0 = slots[0], s1 = slots[1], s2 = slots[2], s3 = slots[3];
          const float a = ((s0.w + s1.w) + (s2.w + s3.w)) + b8;          // output layer (udf_model.py:102)
          float u, gmul;                                                // udf_model.py:82-88 and its derivative
          if (udf_type == 0) { u = fabsf(a); gmul = (a > 0.f) ? 1.f : (a < 0.f ? -1.f : 0.f); }
          else if (udf_type == 1) { u = a * a; gmul = 2.f * a; }
          else { u = a; gmul = 1.f; }
          m.udf_out[pt] = u / net_scale;
          m.grad_out[pt * 3 + 0] = gmul * ((s0.x + s1.x) + (s2.x + s3.x));
          m.grad_out[pt * 3 + 1] = gmul * ((s0.y + s1.y) + (s2.y + s3.y));
          m.grad_out[pt * 3 + 2] = gmul * ((s0.z + s1.z) + (s2.z + s3.z));
        }
        named_bar_sync(1 + q, 128);
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
  if (m.dbg_clk && threadIdx.x == 0) {                      // load-balance diagnostic: dbg_clk[256 + 2 b + {0,1}]
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    m.dbg_clk[256 + 2 * blockIdx.x] = clock64() - t_block0;
    m.dbg_clk[256 + 2 * blockIdx.x + 1] = (long long)smid;
  }
}

static int g_flags = kFlagDynamic | kFlagRolled | kFlagRolledEpi;   // (28) emap_set_option("rg_flags", bits); both measured (B200): dynamic tiles 5.24 vs 6.05 ms, rolled issuer 6.35 vs 7.26 ms with the training stash
int set_flags(int v) { g_flags = v; return 0; }

template <int NTERMS, typename T, int ROLL>
static int launch_r(const Args& a_in, size_t scratch_bytes, cudaStream_t stream);

template <int NTERMS, typename T>
static int launch(const Args& a_in, size_t scratch_bytes, cudaStream_t stream) {
  if (g_flags & kFlagRolledEpi) return launch_r<NTERMS, T, 2>(a_in, scratch_bytes, stream);
  if (g_flags & kFlagRolled) return launch_r<NTERMS, T, 1>(a_in, scratch_bytes, stream);
  return launch_r<NTERMS, T, 0>(a_in, scratch_bytes, stream);
}

template <int NTERMS, typename T, int ROLL>
static int launch_r(const Args& a_in, size_t scratch_bytes, cudaStream_t stream) {
  Args a = a_in;
  a.flags = g_flags;
  const long long tiles = (a.m.P + 127) / 128;
  if (tiles > 0x7fffffffLL) return set_error("too many points");
  a.m.num_tiles = (int)tiles;
  int grid = sm_count();
  if (tiles < grid) grid = (int)tiles;
  a.m.iters = (int)((tiles + grid - 1) / grid);
  const size_t sigma_bytes = (size_t)grid * kSigmaWordsPerCta * sizeof(uint32_t);
  if (scratch_bytes < sigma_bytes + (size_t)grid * kPeBytesPerCta + 256)
    return set_error("emap_udf_forward_grad_rev: scratch too small (%zu bytes, need %zu)", scratch_bytes,
                     sigma_bytes + (size_t)grid * kPeBytesPerCta + 256);
  a.pe_scratch = reinterpret_cast<uint8_t*>(a.scratch) + sigma_bytes;      // PE images behind the sigma slices
  a.tile_counter = reinterpret_cast<unsigned int*>(a.pe_scratch + (size_t)grid * kPeBytesPerCta);
  if (a.flags & kFlagDynamic) EMAP_CUDA(cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned int), stream));
  auto kern = mlp_rgrad_kernel<NTERMS, T, ROLL>;
  static bool attr_done = false;   // per template instantiation
  if (!attr_done) {
    EMAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan<NTERMS>::total));
    attr_done = true;
  }
  // Optional: ask the L2 to keep the sigma scratch (rewritten tile after tile) resident instead of letting
  // the weight/stash streams evict it: persisting access-policy window over the slices in use, hit ratio
  // scaled to the carve-out the device grants.  A/B switch -- whether it pays is a measurement (ncu dram bytes).
  const bool persist = (a.flags & kFlagL2Persist) != 0;
  if (persist) {
    int dev = 0, max_persist = 0, max_window = 0;
    EMAP_CUDA(cudaGetDevice(&dev));
    EMAP_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    EMAP_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    size_t bytes = (size_t)grid * kSigmaWordsPerCta * sizeof(uint32_t);
    if (max_persist > 0 && max_window > 0) {
      static bool limit_set = false;
      if (!limit_set) { EMAP_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist)); limit_set = true; }
      if (bytes > (size_t)max_window) bytes = (size_t)max_window;
      cudaStreamAttrValue v;
      memset(&v, 0, sizeof(v));
      v.accessPolicyWindow.base_ptr = a.scratch;
      v.accessPolicyWindow.num_bytes = bytes;
      v.accessPolicyWindow.hitRatio = (bytes <= (size_t)max_persist) ? 1.0f : (float)max_persist / (float)bytes;
      v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
      EMAP_CUDA(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v));
    }
  }
  if (IsFp16<T>::value && a.m.st_u && !(a.flags & kFlagRegStash) &&
      make_stash_map(a.m.stash_map, a.m.st_u, a.m.P, 128)) return 1;
  kern<<<grid, kThreadsRg, Plan<NTERMS>::total, stream>>>(a);
  EMAP_CUDA(cudaGetLastError());
  if (persist) {
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));                       // num_bytes = 0 disables the window for later launches
    EMAP_CUDA(cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &v));
  }
  return 0;
}

int run(const emap_net_desc* net, const void* packed, int precision, const float* pts, const float* rays_o,
        const float* rays_d, const float* z, int32_t n_per_ray, int64_t P, float* udf_out, float* grad_out,
        void* scratch, size_t scratch_bytes, void* st_u0, void* st_u, float* dbg_acc, void* stream);

}  // namespace rg
}  // namespace emap

using namespace emap;

// Test hook on HOST memory (no GPU needed): the kernel's own PE-adjoint contraction (pe_adjoint16) applied to
// 16 adjoints of the slots kbase..kbase+15; g3 is accumulated into.
extern "C" int emap_debug_pe_adjoint(const float* adj16, int kbase, const float* x3, int multires, float* g3) {
  if (!adj16 || !x3 || !g3 || (kbase & 1)) return set_error("emap_debug_pe_adjoint: bad arguments (kbase must be even)");
  float adj[16], x[3] = {x3[0], x3[1], x3[2]}, g[3] = {g3[0], g3[1], g3[2]};
  for (int i = 0; i < 16; ++i) adj[i] = adj16[i];
  rg::pe_adjoint16(adj, kbase, x, multires, g);
  g3[0] = g[0]; g3[1] = g[1]; g3[2] = g[2];
  return 0;
}

extern "C" size_t emap_rgrad_scratch_bytes(void) {
  return (size_t)sm_count() * (rg::kSigmaWordsPerCta * sizeof(uint32_t) + rg::kPeBytesPerCta) + 256;
}

extern "C" int emap_udf_forward_grad_rev(const emap_net_desc* net, const void* packed, int precision,
                                         const float* pts, const float* rays_o, const float* rays_d,
                                         const float* z, int32_t n_per_ray, int64_t P, float* udf_out,
                                         float* grad_out, void* scratch, size_t scratch_bytes, void* st_u0,
                                         void* st_u, void* stream) {
  return rg::run(net, packed, precision, pts, rays_o, rays_d, z, n_per_ray, P, udf_out, grad_out, scratch,
                 scratch_bytes, st_u0, st_u, nullptr, stream);
}

// Test hook: the same launch, additionally dumping the accumulators of tile 0 after every MMA step,
// dbg_acc[16][128][256] floats: steps 0..7 = W_l h_l (no bias), steps 8..14 = alpha_l W_l for l = 7..1 (the
// adjoint of layer l's input, unsigned seed, de-scaled), step 15 = the 64 PE slots (columns 0..63).
extern "C" int emap_debug_rgrad(const emap_net_desc* net, const void* packed, int precision,
                                const float* pts, const float* rays_o, const float* rays_d,
                                const float* z, int32_t n_per_ray, int64_t P, float* udf_out,
                                float* grad_out, void* scratch, size_t scratch_bytes, float* dbg_acc,
                                void* stream) {
  return rg::run(net, packed, precision, pts, rays_o, rays_d, z, n_per_ray, P, udf_out, grad_out, scratch,
                 scratch_bytes, nullptr, nullptr, dbg_acc, stream);
}

int emap::rg::run(const emap_net_desc* net, const void* packed, int precision, const float* pts,
                  const float* rays_o, const float* rays_d, const float* z, int32_t n_per_ray, int64_t P,
                  float* udf_out, float* grad_out, void* scratch, size_t scratch_bytes, void* st_u0,
                  void* st_u, float* dbg_acc, void* stream) {
  if (check_net(net)) return 1;
  if ((st_u0 == nullptr) != (st_u == nullptr)) return set_error("emap_udf_forward_grad_rev: give both stash pointers or none");
  if (!packed || !udf_out || !grad_out || !scratch) return set_error("emap_udf_forward_grad_rev: NULL pointer");
  if (P <= 0) return set_error("P must be > 0");
  if (!pts) {
    if (!rays_o || !rays_d || !z) return set_error("give either pts or (rays_o, rays_d, z)");
    if (n_per_ray <= 0 || P % n_per_ray) return set_error("P must be a multiple of n_per_ray");
  }
  rg::Args a;
  memset(&a, 0, sizeof(a));
  a.m.packed = (const uint8_t*)packed; a.m.pts = pts; a.m.rays_o = rays_o; a.m.rays_d = rays_d; a.m.z = z;
  a.m.n_per_ray = n_per_ray; a.m.P = P; a.m.udf_out = udf_out; a.m.grad_out = grad_out;
  a.m.dbg_acc = dbg_acc;
  a.m.dbg_clk = dbg_acc ? dbg_clk_buffer() : nullptr;      // timeline stamps only through emap_debug_rgrad
  a.m.dbg_iter = dbg_iter();
  a.m.st_u0 = (__half*)st_u0; a.m.st_u = (__half*)st_u;
  a.scratch = (uint32_t*)scratch;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == EMAP_PREC_FP32X3) {
    if (net->elem_type == 0) return rg::launch<3, __half>(a, scratch_bytes, st);
    return rg::launch<3, __nv_bfloat16>(a, scratch_bytes, st);
  } else if (precision == EMAP_PREC_HALF) {
    if (net->elem_type == 0) return rg::launch<1, __half>(a, scratch_bytes, st);
    return rg::launch<1, __nv_bfloat16>(a, scratch_bytes, st);
  }
  return set_error("precision must be EMAP_PREC_FP32X3 (3) or EMAP_PREC_HALF (1)");
}
