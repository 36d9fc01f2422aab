// Common device helpers for the sm_100a kernels of emap_b200:
// mbarrier / bulk-copy (TMA engine) / tcgen05 (UMMA + TMEM) PTX wrappers, descriptor builders,
// and the network geometry shared by the pack kernel, the MLP kernels and the C-ABI.
//
// Everything here is written against the PTX ISA directly (no CUTLASS/CuTe dependency).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace emap {

// ----------------------------------------------------------------------------------------------
// Network geometry.  The reference topology is fixed by confs/*.conf (udf_network block):
//   d_in=3, d_hidden=256, n_layers=8 (=> 9 Linear), skip_in=[4], d_out=1, multires L in [0,10].
// PE width pe = 3 + 6L <= 63; layer 3 has 256-pe outputs; layer 4 consumes [h4 ; PE]/sqrt(2).
// (reference: src/models/udf_model.py:24-76)
// ----------------------------------------------------------------------------------------------
constexpr int kHidden = 256;
constexpr int kNumLinear = 9;
constexpr int kSkipLayer = 4;
constexpr int kMaxFreq = 10;
constexpr float kSoftplusBeta = 100.0f;       // udf_model.py:78
constexpr float kWeightScale = 16.0f;         // power-of-two pre-scale of packed fp16 weights
constexpr float kInvWeightScale = 1.0f / 16.0f;

// PE column order used INSIDE the kernels (a permutation of the reference order, chosen so that
// the two threads that share a row each own 32 contiguous columns = four 16-byte swizzle groups):
//   col 0..2  : x_c            col 3 : zero pad
//   pair q = 3*j + c (frequency j, axis c), q in [0,30):
//     q < 14 : cols 4+2q (sin), 5+2q (cos)          (thread half 0: cols 0..31)
//     q >= 14: cols 32+2(q-14) (sin), +1 (cos)      (thread half 1: cols 32..63)
// Reference order (embedder.py:26-35): [x(3), sin(f0 x)(3), cos(f0 x)(3), sin(f1 x)(3), ...].
__host__ __device__ inline int pe_col_to_ref(int col, int multires) {
  // returns reference PE index for kernel column `col`, or -1 if the column is padding
  if (col < 3) return col;
  if (col == 3) return -1;
  int q, is_cos;
  if (col < 32) { q = (col - 4) >> 1; is_cos = (col - 4) & 1; }
  else          { q = 14 + ((col - 32) >> 1); is_cos = (col - 32) & 1; }
  int j = q / 3, c = q % 3;
  if (j >= multires) return -1;
  return 3 + 6 * j + 3 * is_cos + c;
}

// PE row order of the REVERSE-mode gradient kernel (mlp_rg.cu): the adjoint of the positional encoding
// comes out of two MMAs -- layer 4's skip input (accumulator columns out3-1+k) and layer 0 (columns k) --
// and both use this order so that one contraction routine serves both.  It is the kernel column order
// above with x_0 moved into the pad slot: slot 0 of the skip layer is the last hidden column out3-1, so
//   k = 0 : not a PE entry (-1)      k = 1, 2 : x_1, x_2      k = 3 : x_0      k >= 4 : as pe_col_to_ref.
__host__ __device__ inline int rg_pe_ref(int k, int multires) {
  if (k <= 0) return -1;
  if (k == 3) return 0;
  if (k < 3) return k;
  return pe_col_to_ref(k, multires);
}

// ----------------------------------------------------------------------------------------------
// Ring items: the weight stream of one tile, in consumption order.  Built on the host by
// build_item_table() (pack.cu), stored in the packed-weights buffer, copied to SMEM by the MLP
// kernel.  One item = one [n_rows x 64] K-major, 128B-swizzled fp16/bf16 image (<= 16 KiB).
// ----------------------------------------------------------------------------------------------
struct __align__(16) RingItem {
  uint32_t gmem_off;     // byte offset of the image inside the packed buffer (16B aligned)
  uint16_t bytes16;      // image bytes / 16
  uint8_t  layer;        // 0..8
  uint8_t  a_chunk;      // A-operand chunk: 0..3 = activation tile chunk, 4 = PE chunk
  uint8_t  n_off8;       // accumulator column offset / 8   (0 or 16)
  uint8_t  n_rows8;      // UMMA N / 8
  uint8_t  part;         // 0 = W_hi image, 1 = W_lo image
  uint8_t  flags;        // see kItem* below
  uint32_t pad;
};
static_assert(sizeof(RingItem) == 16, "RingItem must be 16 bytes");
constexpr uint8_t kItemFirstOfAcc   = 1;   // first MMA into this accumulator half: overwrite (scale_c=0)
constexpr uint8_t kItemLastOfLayer  = 2;   // commit acc_full after this item
constexpr uint8_t kItemWaitA        = 4;   // first use of a_chunk in this layer: wait a_ready[a_chunk]
constexpr uint8_t kItemFirstOfLayer = 8;   // wait acc_empty[layer&1] before issuing
constexpr uint8_t kItemChunk0Done   = 16;  // layer 4: last item reading activation chunk 0 -> commit c0_free
constexpr int kMaxItems = 128;
constexpr int kStageBytes = 16384;

// Header of the packed-weights buffer (device memory, written by emap_wn_fold).
struct PackedHeader {
  uint32_t magic;            // 'EMAP'
  uint32_t multires;
  uint32_t n_items[2];       // [0]: 1-term table, [1]: 3-term table
  uint32_t items_off[2];     // byte offsets of the two RingItem tables
  uint32_t images_off;       // byte offset of the image area (1024B aligned)
  uint32_t images_bytes;
  uint32_t bias100_off;      // float[9][256]: 100*b (layers 0..7), layer 8: b8 at [8][0]
  uint32_t weff_off;         // float W_eff, layer l at weff_layer_off[l] (row-major [out,in])
  uint32_t weff_layer_off[kNumLinear];
  uint32_t out_dim[kNumLinear];
  uint32_t in_dim[kNumLinear];
  uint32_t total_bytes;
  uint32_t elem_type;        // 0 = fp16 images, 1 = bf16 images
  float    scale;            // UDFNetwork.scale
  uint32_t udf_type;         // 0 abs, 1 square, 2 sdf
  uint32_t reserved[6];
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Remote arrive on the same-offset barrier of CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU box.  The production
// build keeps the slow path minimal -- a spin counter and a trap: the diagnostic variant (clock64 + printf of
// the barrier's tag, compile with -DEMAP_BARRIER_DIAG for bring-up) was inlined ~240 times per MLP kernel and
// made up 45 % of its SASS (ncu, round 2: instruction-cache hit rate 76 %, 15 % of the warp samples without
// an instruction to issue).  try_wait suspends the thread for a hardware time slice per call, so the spin
// limit corresponds to seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0, int info = -1) {
  if (mbar_try_wait(bar, parity)) return;
#ifdef EMAP_BARRIER_DIAG
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 6000000000LL) {  // ~3-4 s
      if ((threadIdx.x & 31) == 0)
        printf("emap: mbarrier timeout tag=%d info=%d block=%d warp=%d parity=%u\n", tag, info,
               blockIdx.x, threadIdx.x >> 5, parity);
      __nanosleep(20000000);   // let the other roles report before the trap tears the grid down
      __trap();
    }
  }
#else
  (void)tag; (void)info;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
#endif
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// generic <-> async proxy fence over every state space: global data written with ordinary stores is about to be
// read by the TMA engine (cp.async.bulk)
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

// TMA-engine bulk copy global -> shared (1-D), completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// Same, multicast to every CTA in `cta_mask` of the cluster (same CTA-relative dst / barrier offsets).
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src,
                                                   uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (fp16 or bf16 inputs, fp32 accumulate).
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- 2-CTA (cta_group::2) variants: one MMA instruction spans the tensor cores of a CTA pair ----------
// acquire at cluster scope: the barrier is arrived on by the peer CTA (mbar_arrive_cluster)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int tag = 0, int info = -1) {
  if (mbar_try_wait_cluster(bar, parity)) return;
#ifdef EMAP_BARRIER_DIAG
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 6000000000LL) {
      if ((threadIdx.x & 31) == 0)
        printf("emap: mbarrier(cluster) timeout tag=%d info=%d block=%d warp=%d parity=%u\n", tag, info,
               blockIdx.x, threadIdx.x >> 5, parity);
      __nanosleep(20000000);
      __trap();
    }
  }
#else
  (void)tag; (void)info;
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
#endif
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[2 x 128 rows, one half per CTA] * B[N rows, one half per CTA]^T; issued by
// the leader CTA only, descriptors are CTA-relative (same offsets in both CTAs).
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// 32 lanes x 32 columns of fp32 accumulators: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_32x32b_x1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

// ---- UMMA descriptors -----------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 64 16-bit elements
// (128 B); 8 rows form one 1024 B swizzle atom; SBO = 1024 B between 8-row groups; LBO unused.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;              // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;      // SBO
  d |= static_cast<uint64_t>(1) << 46;              // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;              // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: D=fp32, A/B = fp16 (fmt 0) or bf16 (fmt 1), both K-major.
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N, int fmt) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// Byte offset of element (row, col) inside one [rows x 64] 16-bit K-major SW128 image.
__host__ __device__ inline uint32_t sw128_offset(int row, int col) {
  return static_cast<uint32_t>(row) * 128u + ((((col >> 3) ^ (row & 7)) & 7) << 4) + ((col & 7) << 1);
}

// ---- 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): one full 32-byte sector per lane
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// streaming variant: written once, read by a later kernel -- do not let it push the L2-resident scratch out
__device__ __forceinline__ void stg256_cs(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

// ---- small math helpers ----------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace emap
