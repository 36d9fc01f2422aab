// K1 / K1g: fused positional-encoding + 9-layer softplus MLP on the 5th-gen tensor cores.
//
// Replaces (reference paths relative to /root/reference):
//   src/models/embedder.py:26-35      Embedder.embed           -> input stage (sincosf, fp32)
//   src/models/udf_model.py:90-110    UDFNetwork.forward       -> MODE_FWD
//   src/models/udf_model.py:121-135   UDFNetwork.gradient      -> MODE_GRAD (forward-mode tangents)
//
// Dataflow (one persistent CTA per SM, 10 warps, warp-specialised):
//   warp 16  : producer -- streams the pre-swizzled weight images (pack.cu) from L2 into a ring of
//              32 KiB stages (one [256 x 64] N x K operand = both N halves of one part) with 1-D bulk
//              copies (TMA engine, SASS UBLKCP), mbarrier tx-count.  The schedule is a fixed function
//              of (layer, K chunk, part): no table, one elected lane, ~20 instructions per copy.
//   warp 17  : MMA issuer -- one elected thread issues tcgen05.mma (M=128, N=256, K=16, kind::f16)
//              with A = activation tile in shared memory (K-major, 128B swizzle), B = weight stage,
//              D = fp32 accumulator in TMEM (two 256-column buffers, ping-pong across layers).
//   warps 0-15: epilogue -- tcgen05.ld the accumulator, softplus (beta=100) in fp32, convert to
//              fp16 (hi [+ lo]) and write the NEXT layer's A tile in place, 64-column chunk by chunk;
//              the MMA of layer l+1 starts on chunk c as soon as it is written, so it overlaps the
//              epilogue of layer l.  Activations never leave the SM.
//
// Precision modes (template NTERMS):
//   3 : "fp32-faithful": every operand is split x = hi + lo (two fp16), and the product is formed as
//       hi*hi + lo*hi + hi*lo with fp32 accumulation (3 MMAs) -> ~2^-22 relative, i.e. fp32-class.
//   1 : one fp16 (or bf16) MMA.
//
// MODE_GRAD: a tile is 32 points x {value, d/dx, d/dy, d/dz} rows (row = 4*p + type inside each
// 32-lane TMEM quarter).  Tangent rows carry J_gamma e_c through the same GEMMs; their epilogue is
// multiplication by softplus'(a) = sigmoid(100 a) of the value row (exchanged through a warp-private
// shared-memory scratch so that the transcendental work is spread over all 32 lanes).
#include "mlp_dev.cuh"

namespace emap {

namespace rg { int set_flags(int v); }   // mlp_rg.cu
namespace dw { int set_desc_strides(int which, int v); }   // mlp_dw.cu
namespace rev { int set_dynamic(int v); }                  // mlp_rev.cu

// The tangent forward with TMA-staged stash rows: MODE 3, single fp16 MMA, CL_ = 1 (CL_ = 3 keeps the
// register-staged round-1 form as A/B switch).  It runs a 19th warp that owns the stash traffic.
template <int NTERMS, int MODE, typename T, int CL_> struct TmaStash {
  static constexpr bool value = (MODE == 3 && NTERMS == 1 && IsFp16<T>::value && CL_ == 1);
};
template <int NTERMS, int MODE, typename T, int CL_> struct ThreadsOf {
  static constexpr int value = kThreads + (TmaStash<NTERMS, MODE, T, CL_>::value ? 32 : 0);
};

template <int NTERMS, int MODE, typename T, int CL_>
__global__ void __launch_bounds__(ThreadsOf<NTERMS, MODE, T, CL_>::value, 1) mlp_kernel(const __grid_constant__ MlpArgs args) {
  // CL = 1|2|4: weight-stream multicast width (cta_group::1).  CL = -2: PAIR mode -- clusters of two CTAs
  // driven by ONE MMA issuer with tcgen05.mma.cta_group::2 (M = 256: each CTA's 128-row tile, N = 256 split
  // as 128 weight rows per CTA): every SM stages and reads only half of each weight operand.
  // CL_ = 1 (default) runs the issuer's layer loop ROLLED (schedule computed at run time); CL_ = 3 is the same
  // kernel with the loop unrolled (the round-1 form, kept as A/B switch): the unrolled issuer is ~half of the
  // kernel's SASS and is executed by one thread -- it only costs instruction-cache capacity that the 16 epilogue
  // warps need (measured: K1 forward 2.70 vs 3.04 ms, K1r 6.35 vs 7.26 ms in training mode; profiles/r02_*).
  constexpr bool ROLLI = (CL_ == 1);
  constexpr int CL = (CL_ == 3) ? 1 : CL_;
  constexpr bool PAIR = (CL == -2);
  constexpr int CLW = PAIR ? 2 : CL;             // cluster width
  constexpr bool kTmaC = TmaStash<NTERMS, MODE, T, CL_>::value;
  using Plan = SmemPlan<NTERMS, MODE, PAIR, kTmaC>;
  using MI = ModeInfo<MODE>;
  constexpr int kStages = Plan::kStages;
  constexpr int kStageB = Plan::kStageBytesP;
  constexpr int kArr = PAIR ? 2 : 1;             // epilogue arrivals per barrier: both CTAs of a pair
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(args.packed);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int multires = (int)hdr->multires;
  const float net_scale = hdr->scale;
  const int udf_type = (int)hdr->udf_type;
  const float* bias100 = reinterpret_cast<const float*>(args.packed + hdr->bias100_off);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Plan::bars);
  uint64_t* full = bars;                  // [kStages]
  uint64_t* empty = bars + 8;             // [kStages]
  uint64_t* a_ready = bars + 16;          // [5]  (index 4 = PE written into chunk 0)
  uint64_t* acc_full = bars + 36;         // [2 buffers][2 N halves]: columns [0,128) / [128,256) complete
  uint64_t* acc_empty = bars + 23;        // [2]
  uint64_t* c0_free = bars + 25;          // layer 4 has consumed chunk 0 -> PE may be regenerated there
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);
  uint64_t* peer_full = bars + 28;        // [kStages] PAIR, leader only: the peer's stage has landed
  // Tile schedule of the single-CTA configuration (as in mlp_rg.cu): the tile of iteration k is published in
  // sched_tile[k & 1] by completion k of sched_ready -- one thread of epilogue warp 0, during layer 2 of
  // iteration k-1, when every role has consumed completion k-1 -- either the static round robin or the next
  // value of a global atomic counter (SM-to-SM speed differences of ~10 % otherwise idle the fast SMs at the
  // end).  A tile index >= num_tiles ends every role's loop.  Cluster configurations keep the static loop.
  constexpr bool kSched = (CL == 1);
  // MODE 3, fp16 stashes (the only kind the backward makes): the value rows h_{l+1} the epilogue needs and the
  // tangent rows it produces move through four 16 KiB slots by TMA, issued by a 19th warp -- see the I/O role below
  constexpr bool tma = kTmaC;
  uint64_t* u_full = bars + 40;           // [4] slot c holds h_{l+1} of the coming layer (8 completions per tile)
  uint64_t* out_done = bars + 44;         // [4] slot c holds hdot_{l+1} of the finished layer (8 per tile)
  constexpr int kIoWarp = kEpiWarps + 2;
  uint64_t* sched_ready = bars + 21;
  volatile int* sched_tile = reinterpret_cast<volatile int*>(bars + 22);
  const uint32_t crank = (CLW > 1) ? cluster_ctarank() : 0;

  if (warp == kProducerWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1); mbar_init(&empty[s], PAIR ? 1 : CL); mbar_init(&peer_full[s], 1);
    }
    for (int c = 0; c < 4; ++c) mbar_init(&a_ready[c], kEpiWarps * kArr);
    mbar_init(&a_ready[4], 8 * kArr);
    for (int b = 0; b < 4; ++b) mbar_init(&acc_full[b], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&acc_empty[b], kEpiWarps * kArr);
    mbar_init(c0_free, 1);
    if (kTmaC) { for (int c = 0; c < 4; ++c) { mbar_init(&u_full[c], 1); mbar_init(&out_done[c], kEpiWarps); } }
    if (kSched) { mbar_init(sched_ready, 1); sched_tile[0] = (int)blockIdx.x; }
    fence_barrier_init();
    if (kSched) mbar_arrive(sched_ready);          // completion 0: iteration 0 runs tile blockIdx.x
  }
  if (warp == kMmaWarp) {
    if (PAIR) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CLW > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int kParts = (NTERMS == 3) ? 2 : 1;
  // epilogue -> MMA-issuer arrivals: in PAIR mode the issuer lives in the leader CTA (cluster rank 0)
  auto arrive_issuer = [&](uint64_t* bar) {
    if (!PAIR || crank == 0) mbar_arrive(bar); else mbar_arrive_cluster(bar, 0);
  };

  if (warp == kProducerWarp) {
    // ===================================== producer =====================================
    // The weight stream of a tile is a fixed function of the schedule (pack.cu: layer -> K chunk ->
    // hi/lo part, both N halves of a part adjacent in memory): one 32 KiB bulk copy per part (2 KiB for
    // the N=16 output layer).  Uniform control flow, the copy itself issued by one elected lane.
    const uint32_t rank = crank;
    const uint8_t* img = args.packed + hdr->images_off;
    uint8_t* ring = smem + Plan::ring;
    const bool no_copy = (args.dbg_flags & 2) != 0;
    uint32_t stage = 0, round = 0;
    for (int iter = 0; kSched || iter < args.iters; ++iter) {
      if (kSched) {
        mbar_wait(sched_ready, (uint32_t)iter & 1, 560);
        if (sched_tile[iter & 1] >= args.num_tiles) break;
      }
      uint32_t off = 0;
#pragma unroll 1
      for (int l = 0; l < MI::kLayers; ++l) {
        const int nkc = (l == 0) ? 1 : ((l == kSkipLayer) ? 5 : 4);
        const uint32_t bytes = (l == kNumLinear - 1) ? 2048u : (uint32_t)kRingStageBytes;
#pragma unroll 1
        for (int ip = 0; ip < nkc * 2; ++ip, off += bytes) {
          if (NTERMS == 1 && (ip & 1)) continue;          // single-MMA modes stream the hi images only
          if (round > 0) mbar_wait(&empty[stage], (round - 1) & 1, 100 + (int)stage, l * 16 + ip);
          if (elect_one()) {
            if (no_copy) {
              mbar_arrive(&full[stage]);
            } else {
              uint8_t* dst = ring + stage * kStageB;
              const uint8_t* src = img + off;
              if (PAIR) {
                // this CTA's N half of the operand (rows [128 rank, 128 rank + 128); 8 rows of the N=16 layer)
                mbar_arrive_expect_tx(&full[stage], bytes / 2);
                bulk_g2s(dst, src + rank * (bytes / 2), bytes / 2, &full[stage]);
              } else if (CL == 1) {
                mbar_arrive_expect_tx(&full[stage], bytes);
                bulk_g2s(dst, src, bytes, &full[stage]);
              } else {
                mbar_arrive_expect_tx(&full[stage], bytes);
                const uint32_t slice = bytes / (uint32_t)CLW;
                bulk_g2s_multicast(dst + rank * slice, src + rank * slice, slice, &full[stage],
                                   (uint16_t)((1u << CLW) - 1));
              }
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
        }
      }
    }
  } else if (warp == kMmaWarp && PAIR && crank != 0) {
    // ===================================== peer CTA of a pair: relay =====================
    // The leader issues every MMA; it must know that THIS CTA's half of a weight stage has landed.  Bulk
    // copies can only signal a barrier of their destination CTA, so this otherwise idle warp forwards each
    // `full` completion to the leader's `peer_full`.
    uint32_t stage = 0, round = 0;
    for (int iter = 0; iter < args.iters; ++iter) {
#pragma unroll 1
      for (int l = 0; l < MI::kLayers; ++l) {
        const int nparts = ((l == 0) ? 1 : ((l == kSkipLayer) ? 5 : 4)) * kParts;
#pragma unroll 1
        for (int ip = 0; ip < nparts; ++ip) {
          mbar_wait(&full[stage], round & 1, 450 + (int)stage, l * 16 + ip);
          if (lane == 0) mbar_arrive_cluster(&peer_full[stage], 0);
          __syncwarp();
          if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer ===================================
    // One warp, uniform control flow, the instructions themselves issued by one elected lane (so
    // descriptors live in uniform registers).  Fixed schedule: layer -> K chunk -> hi/lo part; one ring
    // stage = one part = the B operand of N=256 MMAs.  (Measured before this structure: a table-driven
    // issuer under `lane==0` paid ~95 clk per tcgen05.mma and ~600 clk of loop overhead per item; 16 KiB
    // stages consumed in pairs left only two parts in flight and made the producer the bottleneck.)
    const uint32_t a_hi_addr = smem_u32(smem + Plan::a_hi);
    const uint32_t a_lo_addr = smem_u32(smem + Plan::a_lo);
    const uint32_t ring_addr = smem_u32(smem + Plan::ring);
    const uint32_t idesc256 = make_idesc_f16(PAIR ? 256 : 128, 256, Elem<T>::fmt);
    const uint32_t idesc16 = make_idesc_f16(PAIR ? 256 : 128, 16, Elem<T>::fmt);
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
      if (PAIR) umma_f16_2sm(d, a, b, idesc, acc); else umma_f16(d, a, b, idesc, acc);
    };
    auto commit = [&](uint64_t* bar, bool to_cluster) {
      if (PAIR) umma_commit_2sm_multicast(bar, (uint16_t)3);
      else if (to_cluster && CL > 1) umma_commit_multicast(bar, (uint16_t)((1u << CLW) - 1));
      else umma_commit(bar);
    };
    const bool no_mma = (args.dbg_flags & 1) != 0;
    uint32_t stage = 0, round = 0;
    for (int iter = 0; kSched || iter < args.iters; ++iter) {
      if (kSched) {
        mbar_wait(sched_ready, (uint32_t)iter & 1, 561);
        if (sched_tile[iter & 1] >= args.num_tiles) break;
      }
#pragma unroll (ROLLI ? 1 : MI::kLayers)
      for (int l = 0; l < MI::kLayers; ++l) {
        const int buf = l & 1;
        const bool stamp = args.dbg_clk && blockIdx.x == 0 && iter == args.dbg_iter && lane == 0;
        if (stamp) args.dbg_clk[72 + l * 8 + 0] = clock64();
        {
          const uint32_t started = (uint32_t)iter * (buf ? 4u : (uint32_t)MI::kUses0) + (uint32_t)(l >> 1);
          if (started > 0) {
            if (PAIR) mbar_wait_cluster(&acc_empty[buf], (started - 1) & 1, 200 + buf, l);
            else mbar_wait(&acc_empty[buf], (started - 1) & 1, 200 + buf, l);
          }
        }
        if (stamp) args.dbg_clk[72 + l * 8 + 1] = clock64();
        const int nkc = (l == 0) ? 1 : ((l == kSkipLayer) ? 5 : 4);
        const bool last = (l == kNumLinear - 1);
        const uint32_t idesc = last ? idesc16 : idesc256;
        const uint32_t d = tmem_base + (uint32_t)buf * 256u;
        // N-split of the last K chunk (below).  Not for layers 0 and 4: their last K chunk is the PE chunk,
        // which lives in activation chunk 0 -- the epilogue of accumulator half 0 would overwrite it under
        // the still-running MMAs of half 1.
        const bool split_tail = (NTERMS == 3) && !PAIR && !last && l != 0 && l != kSkipLayer;   // single-MMA modes: measured no gain
#pragma unroll
        for (int ic = 0; ic < nkc; ++ic) {
          const int c = (l == 0) ? 4 : ((ic < 4) ? ic : 4);
          {
            const uint32_t uses = (c == 4) ? (uint32_t)iter * 2u + (l == kSkipLayer ? 1u : 0u)
                                           : (uint32_t)iter * (uint32_t)MI::kAPerTile + (uint32_t)(l - 1);
            if (PAIR) mbar_wait_cluster(&a_ready[c], uses & 1, 300 + c, l);
            else mbar_wait(&a_ready[c], uses & 1, 300 + c, l);
          }
          tc_fence_after();
          if (stamp && ic < 4) args.dbg_clk[72 + l * 8 + 2 + ic] = clock64();
          const uint32_t coff = (c == 4) ? 0u : (uint32_t)c * kChunkBytes;   // PE lives in chunk 0
          const uint64_t ahi = make_sw128_kmajor_desc(a_hi_addr + coff);
          const uint64_t alo = make_sw128_kmajor_desc(a_lo_addr + coff);
          if (split_tail && ic == nkc - 1) {
            // Last K chunk of the layer, N split in halves: the MMAs into accumulator columns [0,128) are
            // issued (and committed to acc_full[buf][0]) first, so the epilogue starts on chunks 0-1 while
            // the tensor pipe still works on columns [128,256) -- the accumulator tail (12 MMAs in fp32x3
            // mode) is no longer serial with the first-chunk latency of the next layer.
            uint32_t sidx[kParts];
#pragma unroll
            for (int part = 0; part < kParts; ++part) {
              sidx[part] = stage;
              mbar_wait(&full[stage], round & 1, 400 + (int)stage, l * 16 + ic * 2 + part);
              if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
            }
            tc_fence_after();
            const uint32_t idesc128 = make_idesc_f16(128, 128, Elem<T>::fmt);
            if (elect_one()) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t dh = d + (uint32_t)half * 128u;
                const uint64_t b0 = make_sw128_kmajor_desc(ring_addr + sidx[0] * kStageB + half * kStageBytes);
                if (!no_mma) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16(dh, ahi + 2 * k, b0 + 2 * k, idesc128, (ic == 0 && k == 0) ? 0u : 1u);
                  if (NTERMS == 3) {
                    const uint64_t b1 = make_sw128_kmajor_desc(ring_addr + sidx[kParts - 1] * kStageB + half * kStageBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(dh, alo + 2 * k, b0 + 2 * k, idesc128, 1u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(dh, ahi + 2 * k, b1 + 2 * k, idesc128, 1u);
                  }
                }
                commit(&acc_full[buf * 2 + half], false);
              }
#pragma unroll
              for (int part = 0; part < kParts; ++part) commit(&empty[sidx[part]], true);
            }
            __syncwarp();
            continue;
          }
#pragma unroll
          for (int part = 0; part < kParts; ++part) {
            const bool st2 = stamp && l == 2 && ic == 1;
            if (st2) args.dbg_clk[144 + part * 4 + 0] = clock64();
            mbar_wait(&full[stage], round & 1, 400 + (int)stage, l * 16 + ic * 2 + part);
            if (PAIR) mbar_wait_cluster(&peer_full[stage], round & 1, 420 + (int)stage, l * 16 + ic * 2 + part);
            tc_fence_after();
            if (st2) args.dbg_clk[144 + part * 4 + 1] = clock64();
            const uint64_t bdesc = make_sw128_kmajor_desc(ring_addr + stage * kStageB);
            if (elect_one()) {
              if (!no_mma) {
                if (part == 0) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    mma(d, ahi + 2 * k, bdesc + 2 * k, idesc, (ic == 0 && k == 0) ? 0u : 1u);
                  if (NTERMS == 3) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma(d, alo + 2 * k, bdesc + 2 * k, idesc, 1u);
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k) mma(d, ahi + 2 * k, bdesc + 2 * k, idesc, 1u);
                }
              }
              commit(&empty[stage], true);
            }
            __syncwarp();
            if (st2) args.dbg_clk[144 + part * 4 + 2] = clock64();
            if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
          }
          if (l == kSkipLayer && ic == 0) { if (elect_one()) commit(c0_free, false); __syncwarp(); }
        }
        if (!split_tail) {
          if (elect_one()) { commit(&acc_full[buf * 2], false); if (NTERMS == 3) commit(&acc_full[buf * 2 + 1], false); }
          __syncwarp();
        }
        if (stamp) args.dbg_clk[72 + l * 8 + 6] = clock64();
      }
    }
  } else if (kTmaC && warp == kIoWarp) {
    // ===================================== stash I/O (MODE 3): one thread ================
    // Slot c (16 KiB: [128 points x 64 columns], 128-byte swizzle) is filled with the value rows h_{l+1} of the
    // tile a whole layer ahead of their use, overwritten in place by the epilogue with the tangent rows
    // hdot_{l+1}, written to the stash by a TMA store, and refilled as soon as the store has read it.
    if (lane == 0) {
      uint8_t* slots = smem + Plan::slots;
      const void* map = args.stash_map;
      auto load_h = [&](int c, int plane, int pt0) {
        mbar_arrive_expect_tx(&u_full[c], kChunkBytes);
        tma_load_3d(slots + c * kChunkBytes, map, c * 64, pt0, plane * 2, &u_full[c]);
      };
      for (int iter = 0;; ++iter) {
        mbar_wait(sched_ready, (uint32_t)iter & 1, 563);
        const int tile = sched_tile[iter & 1];
        if (tile >= args.num_tiles) break;
        const int pt0 = tile * 128;
        if (iter == 0) {
          for (int c = 0; c < 4; ++c) load_h(c, 0, pt0);
        }
        int next_tile = args.num_tiles;
#pragma unroll 1
        for (int l = 0; l < 8; ++l) {
          auto refill = [&](int c) {
            if (l < 7) load_h(c, l + 1, pt0);
            else if (next_tile < args.num_tiles) load_h(c, 0, next_tile * 128);
          };
          if (l == 7) {                                         // the next tile: published during layer 2 of this one
            mbar_wait(sched_ready, (uint32_t)(iter + 1) & 1, 564);
            next_tile = sched_tile[(iter + 1) & 1];
          }
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            mbar_wait(&out_done[c], ((uint32_t)iter * 8u + (uint32_t)l) & 1, 600 + c, l);
            tma_store_3d(slots + c * kChunkBytes, map, c * 64, pt0, l * 2 + 1);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (c > 0) {                                        // one store stays in flight
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              refill(c - 1);
            }
          }
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          refill(3);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store has left before the CTA exits
    }
    __syncwarp();
  } else {
    // ===================================== epilogue warps ================================
    // warp = 4*sub + q: q = TMEM lane quarter; per layer each warp converts two 32-column pieces:
    // phase 0 -> chunk (sub>>1), phase 1 -> chunk 2+(sub>>1); piece (sub&1) of the chunk.  Chunks 0,1 are
    // therefore complete after half of the epilogue and the next layer's MMA starts on them.
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;                 // TMEM lane == A-tile row
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* A_hi = smem + Plan::a_hi;
    uint8_t* A_lo = smem + Plan::a_lo;
    float* sc = reinterpret_cast<float*>(smem + Plan::scratch) + warp * kScratchFloatsPerWarp;   // kOwnScratch only
    const float k1 = kSoftplusBeta * kInvWeightScale;
    const int p8 = lane >> 2, ty = lane & 3;       // MODE_GRAD: point-in-warp, row type
    const float b8 = bias100[8 * kHidden];

    const bool scheduler = kSched && warp == 0 && lane == 0;
    for (int iter = 0; kSched || iter < args.iters; ++iter) {
      long long tile = (long long)blockIdx.x + (long long)iter * gridDim.x;
      int next_tile = 0;
      if (kSched) {
        mbar_wait(sched_ready, (uint32_t)iter & 1, 562);
        tile = (long long)sched_tile[iter & 1];
        if (tile >= args.num_tiles) break;
        if (scheduler) {       // the tile after this one: fetched now (an L2 atomic in dynamic mode), published in layer 2
          const long long nt = args.tile_counter ? (long long)gridDim.x + (long long)atomicAdd(args.tile_counter, 1u)
                                                 : tile + (long long)gridDim.x;
          next_tile = (nt < (long long)args.num_tiles) ? (int)nt : args.num_tiles;
        }
      }
      // ------------------------------------------------ input stage: positional encoding -> chunk 0
      long long pt;
      if (MODE == 0 || MODE == 3) pt = tile * 128 + row;
      else if (MODE == 1) pt = tile * 32 + q * 8 + p8;
      else pt = tile * 64 + q * 16 + (lane >> 1);
      float x[3];
      load_point(args, pt, net_scale, x);
      float gb[3] = {0.f, 0.f, 0.f};
      if (MODE >= 2 && args.gbar) {
        const long long pc = (pt < args.P) ? pt : args.P - 1;
        const float sg = args.bwd_scales ? __ldg(args.bwd_scales) : 1.f;     // power of two (loss scaling)
        gb[0] = sg * args.gbar[pc * 3]; gb[1] = sg * args.gbar[pc * 3 + 1]; gb[2] = sg * args.gbar[pc * 3 + 2];
      }
      if (sub < 2) {
        if (sub == 0) pe_stage<NTERMS, MODE, T, 0>(args, x, multires, lane, row, pt, tile, true, A_hi, A_lo, gb);
        else          pe_stage<NTERMS, MODE, T, 1>(args, x, multires, lane, row, pt, tile, true, A_hi, A_lo, gb);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive_issuer(&a_ready[4]);
      }

      // ------------------------------------------------ hidden layers 0..7
      for (int l = 0; l < 8; ++l) {
        const int buf = l & 1;
        const bool stamp = args.dbg_clk && blockIdx.x == 0 && iter == args.dbg_iter && warp == 0 && lane == 0;
        if (stamp) args.dbg_clk[l * 8 + 0] = clock64();
        const uint32_t acc_par = ((uint32_t)iter * (buf ? 4u : (uint32_t)MI::kUses0) + (uint32_t)(l >> 1)) & 1;
        // MODE 3: this row's h_{l+1} (value row of the stash, written by the training forward) for the
        // thread's 16 columns of all four chunks -- independent of the MMA, fetched before waiting for it
        uint32_t hw[(MODE == 3) ? 4 : 1][8];
        if (MODE == 3 && !tma) {
          const long long pc = (pt < args.P) ? pt : args.P - 1;
          const __half* hp = args.st_u + (size_t)l * 2 * (size_t)args.P * 256 + (size_t)pc * 256 + sub * 16;
          if (args.dbg_flags & 32) {            // (dbg 32: timing experiment without the stash loads)
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
              for (int k = 0; k < 8; ++k) hw[c4][k] = 0x3c003c00u;
          } else {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) ldg256(hp + c4 * 64, hw[c4]);
          }
          // (an L2 prefetch of the next layer's rows from here was measured: 4.30 vs 4.04 ms -- slower, removed)
        }
        mbar_wait(&acc_full[buf * 2], acc_par, 500 + buf, l);
        tc_fence_after();
        if (stamp) args.dbg_clk[l * 8 + 1] = clock64();
        if (l == 2 && scheduler) {      // all 16 warps are past layer 1 of this tile, i.e. past its schedule wait
          sched_tile[(iter + 1) & 1] = next_tile;
          mbar_arrive(sched_ready);
        }
        const float* bl = bias100 + l * kHidden;
#pragma unroll (MODE == 3 ? 4 : 1)
        for (int chunk = 0; chunk < 4; ++chunk) {
          // all 16 warps convert the same 64-column chunk (16 columns each), so chunk c of the next
          // layer's A tile is complete after (c+1)/4 of the epilogue and its MMAs start then
          // (software-pipelining the tcgen05.ld one chunk ahead was measured: slower -- register pressure)
          uint8_t* dst_hi = A_hi + chunk * kChunkBytes;
          uint8_t* dst_lo = A_lo + chunk * kChunkBytes;
          const int col0 = chunk * 64 + sub * 16;
          if (NTERMS == 3 && chunk == 2) { mbar_wait(&acc_full[buf * 2 + 1], acc_par, 505 + buf, l); tc_fence_after(); }
          // the bias words this lane needs, issued before the accumulator load so that their L1/L2
          // latency hides under the tcgen05.ld wait (ncu: 8 % of the dual kernel's samples sat on it)
          //   MODE 0: 16 columns; MODE 1: the lane's 4 value columns; MODE 2: the lane's 8 columns
          constexpr int kBiasVec = (MODE == 0) ? 4 : ((MODE == 1 || MODE == 3) ? 1 : 2);   // MODE 3: unused
          float4 bv[kBiasVec];
          {
            const int bofs = (MODE == 0 || MODE == 3) ? 0 : ((MODE == 1) ? 4 * ty : 8 * (lane & 1));
#pragma unroll
            for (int i = 0; i < kBiasVec; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(bl + col0 + bofs) + i);
          }
          uint32_t r[16];
          tmem_ld_32x32b_x16(lane_taddr + (uint32_t)(buf * 256 + col0), r);
          tmem_wait_ld();
          if (stamp && chunk < 2) args.dbg_clk[l * 8 + 2 + 3 * chunk] = clock64();
          if (args.dbg_acc && tile == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              args.dbg_acc[((size_t)l * 128 + row) * 256 + col0 + k] = __uint_as_float(r[k]) * kInvWeightScale;
          }
          // MODE 2 / 3: this thread's 32 bytes of the stash row.  Stored AFTER the chunk's hand-off: the
          // fence.proxy.async in front of the hand-off compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and a global
          // store issued before it makes every hand-off wait for an L2 round trip (ncu: 8 % of the tangent
          // forward's samples sat on that fence)
          uint32_t pu[8];
          __half* st_dst = nullptr;
          if (args.dbg_flags & 4) {
            // timing experiment: no epilogue math / stores
          } else if (MODE == 2) {
            // dual rows (lane pair = value, tangent).  The pair splits the transcendental work: the value
            // lane runs softplus/sigmoid on columns 0-7 of the value accumulator, the tangent lane on
            // columns 8-15 (fetched by shuffle); then the value lane receives h[8..15] and the tangent lane
            // sigma[0..7].  Value lane keeps h, tangent lane keeps sigma * adot.  Everything is also
            // stashed (fp16, row-major) for the reverse sweep.
            const int t2 = lane & 1;
            const bool okp = (tile < args.num_tiles) && (pt < args.P);
            const long long rowg = (t2 ? args.P : 0) + pt;
            const float4 bA = bv[0], bB = bv[kBiasVec - 1];
            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
            float hm[8], sm[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t up = __shfl_xor_sync(0xffffffffu, r[8 + j], 1);   // value lane's acc[8+j]
              const float aval = __uint_as_float(t2 ? up : r[j]);
              hm[j] = softplus100<true>(fmaf(aval, k1, bb[j]), sm[j]);
            }
            float outv[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float y = __shfl_xor_sync(0xffffffffu, t2 ? hm[j] : sm[j], 1);   // -> h[8+j] | sigma[j]
              const float ad_lo = __uint_as_float(r[j]) * kInvWeightScale;
              const float ad_hi = __uint_as_float(r[8 + j]) * kInvWeightScale;
              outv[j] = t2 ? y * ad_lo : hm[j];
              outv[8 + j] = t2 ? sm[j] * ad_hi : y;
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float v8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v8[j] = outv[g * 8 + j];
              if (l < 7) store_group<NTERMS, T>(dst_hi, dst_lo, row, sub * 2 + g, v8);
#pragma unroll
              for (int j = 0; j < 4; ++j) pu[g * 4 + j] = Elem<__half>::pack2(v8[2 * j], v8[2 * j + 1]);   // 16 columns = 32 B per stash row
            }
            if (okp && !(args.dbg_flags & 8))
              st_dst = args.st_u + (size_t)l * 2 * (size_t)args.P * 256 + (size_t)rowg * 256 + col0;
          } else if (MODE == 3) {
            // tangent rows only: hdot_{l+1} = softplus'(a_l) . adot_l with softplus'(a_l) = 1 - exp(-100 h_{l+1})
            // recovered from the value row of the stash (as the reverse sweep does); next layer's A tile +
            // the tangent row of the stash.
            const bool okp = (tile < args.num_tiles) && (pt < args.P);
            // TMA path: this row's 16 columns of h_{l+1} are two 16-byte units of the slot (row = point)
            uint8_t* srow = smem + Plan::slots + chunk * kChunkBytes + (uint32_t)row * 128u;
            if (kTmaC) {
              mbar_wait(&u_full[chunk], ((uint32_t)iter * 8u + (uint32_t)l) & 1, 610 + chunk, l);
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                const uint4 v = *reinterpret_cast<const uint4*>(srow + ((((uint32_t)(sub * 2 + g)) ^ (uint32_t)(row & 7)) << 4));
                hw[chunk][g * 4] = v.x; hw[chunk][g * 4 + 1] = v.y; hw[chunk][g * 4 + 2] = v.z; hw[chunk][g * 4 + 3] = v.w;
              }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float v8[8];
#pragma unroll
              for (int w2 = 0; w2 < 4; ++w2) {
                const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&hw[chunk][g * 4 + w2]));
                // sigma / 16: the operand pre-scale's 1/16 folded in (a power of two: rounds exactly as before)
                const float s0 = fmaf(-__expf(-kSoftplusBeta * h2.x), kInvWeightScale, kInvWeightScale);
                const float s1 = fmaf(-__expf(-kSoftplusBeta * h2.y), kInvWeightScale, kInvWeightScale);
                v8[2 * w2] = s0 * __uint_as_float(r[g * 8 + 2 * w2]);
                v8[2 * w2 + 1] = s1 * __uint_as_float(r[g * 8 + 2 * w2 + 1]);
              }
              if (l < 7) store_group<NTERMS, T>(dst_hi, dst_lo, row, sub * 2 + g, v8);
#pragma unroll
              for (int j = 0; j < 4; ++j) pu[g * 4 + j] = Elem<__half>::pack2(v8[2 * j], v8[2 * j + 1]);
              if (kTmaC)                           // the stash row: in place of what was read
                *reinterpret_cast<uint4*>(srow + ((((uint32_t)(sub * 2 + g)) ^ (uint32_t)(row & 7)) << 4)) =
                    make_uint4(pu[g * 4], pu[g * 4 + 1], pu[g * 4 + 2], pu[g * 4 + 3]);
            }
            if (okp && !(args.dbg_flags & 8) && !tma)      // (dbg 8: timing experiment without the stash stores)
              st_dst = args.st_u + (size_t)l * 2 * (size_t)args.P * 256 + ((size_t)args.P + (size_t)pt) * 256 + col0;
          } else if (MODE == 0) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const float4 bA = bv[(2 * g) % kBiasVec], bB = bv[(2 * g + 1) % kBiasVec];
              const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
              float h[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float dummy;
                h[j] = softplus100<false>(fmaf(__uint_as_float(r[g * 8 + j]), k1, bb[j]), dummy);
              }
              store_group<NTERMS, T>(dst_hi, dst_lo, row, sub * 2 + g, h);
            }
          } else {
            // Exchange slots of point p8: four 16-byte slots (columns 4i..4i+3).  Single-MMA modes have a
            // warp-private scratch; in fp32x3 mode (no shared memory to spare) the slots alias this warp's
            // own 64 bytes of one tangent row (hi/lo x two swizzle groups) of the DESTINATION chunk: the
            // MMAs that read it are complete (acc_full), and phase C overwrites it last -- slots
            // (2g, 2g+1) are exactly the hi/lo group g of that row, read by all lanes before the lane that
            // owns the row stores there.
            const int r1 = q * 32 + 4 * p8 + 1 + ((p8 >> 1) % 3);   // (r1 & 7) spread over the 8 points: banks
            auto slot = [&](int i) -> float4* {
              if (Plan::kOwnScratch) return reinterpret_cast<float4*>(sc + p8 * 20 + 4 * i);
              uint8_t* base = (i & 1) ? dst_lo : dst_hi;
              return reinterpret_cast<float4*>(base + (uint32_t)r1 * 128u +
                                               (uint32_t)((((sub * 2 + (i >> 1)) ^ (r1 & 7)) & 7) << 4));
            };
            // phase A: value lanes publish their raw accumulators (16 columns)
            if (ty == 0) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                *slot(i) = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                       __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            }
            __syncwarp();
            // phase B: every lane does 4 columns of its point's value row
            {
              float4* my = slot(ty);
              const float4 vA = *my;
              const float va[4] = {vA.x, vA.y, vA.z, vA.w};
              const float4 bA = bv[0];
              const float bb[4] = {bA.x, bA.y, bA.z, bA.w};
              float h[4], sg[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float sgm;
                h[j] = softplus100<true>(fmaf(va[j], k1, bb[j]), sgm);
                sg[j] = sgm * kInvWeightScale;   // tangent accumulators carry the weight pre-scale
              }
              store_half_group<NTERMS, T>(dst_hi, dst_lo, q * 32 + 4 * p8, sub * 2 + (ty >> 1), ty & 1, h);
              *my = make_float4(sg[0], sg[1], sg[2], sg[3]);
            }
            __syncwarp();
            // phase C: tangent rows: d h = sigmoid(100 a) * d a
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const float4 sA = *slot(2 * g);
              const float4 sB = *slot(2 * g + 1);
              __syncwarp();                       // all lanes have read the slots before any row store
              if (ty != 0) {
                const float ss[8] = {sA.x, sA.y, sA.z, sA.w, sB.x, sB.y, sB.z, sB.w};
                float tv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) tv[j] = ss[j] * __uint_as_float(r[g * 8 + j]);
                store_group<NTERMS, T>(dst_hi, dst_lo, row, sub * 2 + g, tv);
              }
            }
            __syncwarp();
          }
          if (stamp && chunk < 2) args.dbg_clk[l * 8 + 3 + 3 * chunk] = clock64();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(MODE >= 2 && l == 7)) arrive_issuer(&a_ready[chunk]);
          if (kTmaC && lane == 0) mbar_arrive(&out_done[chunk]);
          if (stamp && chunk < 2) args.dbg_clk[l * 8 + 4 + 3 * chunk] = clock64();
          if ((MODE == 2 || MODE == 3) && st_dst) stg256(st_dst, pu);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_issuer(&acc_empty[buf]);

        if (l == kSkipLayer - 1 && sub < 2) {
          // skip connection: layer 4 = [h4 ; PE]/sqrt2.  Once its MMAs on chunk 0 are done, regenerate
          // the PE there (same code, same inputs as at tile start) as the 5th K chunk of layer 4.
          mbar_wait(c0_free, (uint32_t)iter & 1, 520);
          if (sub == 0) pe_stage<NTERMS, MODE, T, 0>(args, x, multires, lane, row, pt, tile, false, A_hi, A_lo, gb);
          else          pe_stage<NTERMS, MODE, T, 1>(args, x, multires, lane, row, pt, tile, false, A_hi, A_lo, gb);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) arrive_issuer(&a_ready[4]);
        }
      }

      // ------------------------------------------------ output layer (layer 8, accumulator buf 0, col 0)
      if (MODE >= 2) continue;        // dual / tangent forward stop at layer 7 (the output layer is pulled back separately)
      mbar_wait(&acc_full[0], ((uint32_t)iter * 5u + 4u) & 1, 510);
      tc_fence_after();
      if (sub == 0) {
        const float accv = __uint_as_float(tmem_ld_32x32b_x1(lane_taddr)) * kInvWeightScale;
        tmem_wait_ld();
        if (args.dbg_acc && tile == 0) args.dbg_acc[((size_t)8 * 128 + row) * 256] = accv;
        const bool ok = (tile < args.num_tiles) && (pt < args.P);
        if (MODE == 0) {
          const float a = accv + b8;
          float u = (udf_type == 0) ? fabsf(a) : (udf_type == 1 ? a * a : a);
          if (ok) args.udf_out[pt] = u / net_scale;
        } else {
          const float a_own = accv + b8;
          const float a = __shfl_sync(0xffffffffu, a_own, lane & ~3);
          if (ty == 0) {
            float u = (udf_type == 0) ? fabsf(a) : (udf_type == 1 ? a * a : a);
            if (ok) args.udf_out[pt] = u / net_scale;
          } else {
            float gmul = 1.f;
            if (udf_type == 0) gmul = (a > 0.f) ? 1.f : (a < 0.f ? -1.f : 0.f);
            else if (udf_type == 1) gmul = 2.f * a;
            if (ok) args.grad_out[pt * 3 + (ty - 1)] = gmul * accv;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_issuer(&acc_empty[0]);
    }
  }

  // ---- teardown
  tc_fence_before();
  if (CLW > 1) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) { if (PAIR) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

static long long* g_dbg_clk = nullptr;   // emap_debug_set_clk_buffer
long long* dbg_clk_buffer() { return g_dbg_clk; }   // (mlp_rg.cu's debug entry stamps into the same buffer)
static int g_dbg_flags = 0;   // timing experiments (emap_set_option("dbg", flags)); 0 in production
static int g_dynamic_tiles = 1;   // emap_set_option("dynamic_tiles", 0): static round-robin tiles (A/B switch)
static int g_dbg_iter = 1;    // which tile iteration of block 0 the clock64 timelines stamp (emap_set_option("dbg_iter"))
int dbg_iter() { return g_dbg_iter; }

// ---------------------------------------------------------------------------------------------
template <int NTERMS, int MODE, typename T, int CL_>
static int launch(const MlpArgs& a_in, cudaStream_t stream) {
  constexpr int CL = (CL_ == 3) ? 1 : CL_;       // 3 = width 1 with the UNROLLED issuer loop (A/B switch)
  constexpr bool PAIR = (CL == -2);
  constexpr int CLW = PAIR ? 2 : CL;
  using Plan = SmemPlan<NTERMS, MODE, PAIR, TmaStash<NTERMS, MODE, T, CL_>::value>;
  MlpArgs a = a_in;
  const int pts_per_tile = ModeInfo<MODE>::kPtsPerTile;
  const long long tiles = (a.P + pts_per_tile - 1) / pts_per_tile;
  if (tiles > 0x7fffffffLL) return set_error("too many points");
  a.num_tiles = (int)tiles;
  a.dbg_flags = g_dbg_flags;
  a.dbg_iter = g_dbg_iter;
  a.tile_counter = (CL == 1 && g_dynamic_tiles) ? tile_counter(stream) : nullptr;
  int grid = sm_count();
  grid = grid / CLW * CLW;
  if (tiles < grid) grid = (int)((tiles + CLW - 1) / CLW * CLW);
  a.iters = (int)((tiles + grid - 1) / grid);
  auto kern = mlp_kernel<NTERMS, MODE, T, CL_>;
  static bool attr_done = false;   // per template instantiation
  if (!attr_done) {
    EMAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::total));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(ThreadsOf<NTERMS, MODE, T, CL_>::value);
  cfg.dynamicSmemBytes = Plan::total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLW;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EMAP_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  return 0;
}

// weight-stream organisation of the MLP kernels (emap_set_option("cluster", v)):
//   1 = every CTA streams its own copy (default), 2 = multicast pairs (cta_group::1),
//   -2 = CTA pairs driven by one cta_group::2 issuer (fp16 operand images only).
static int g_cluster = 1;
static int g_tan_tma = 1;     // emap_set_option("tan_tma", 0): tangent forward with register-staged stash rows

template <int MODE>
static int dispatch(const emap_net_desc* net, int precision, const MlpArgs& a, cudaStream_t st) {
  const int cl = g_cluster;
#define EMAP_LAUNCH(NT, TT)                                              \
  do {                                                                   \
    if (cl == 2) return launch<NT, MODE, TT, 2>(a, st);                  \
    if (cl == 3) return launch<NT, MODE, TT, 3>(a, st);                  \
    return launch<NT, MODE, TT, 1>(a, st);                               \
  } while (0)
  if (precision == EMAP_PREC_FP32X3) {
    if (net->elem_type == 0) { if (cl == -2) return launch<3, MODE, __half, -2>(a, st); EMAP_LAUNCH(3, __half); }
    else EMAP_LAUNCH(3, __nv_bfloat16);
  } else if (precision == EMAP_PREC_HALF) {
    if (net->elem_type == 0) { if (cl == -2) return launch<1, MODE, __half, -2>(a, st); EMAP_LAUNCH(1, __half); }
    else EMAP_LAUNCH(1, __nv_bfloat16);
  }
#undef EMAP_LAUNCH
  return set_error("precision must be EMAP_PREC_FP32X3 (3) or EMAP_PREC_HALF (1)");
}

static int check_points(const float* pts, const float* rays_o, const float* rays_d, const float* z,
                        int n_per_ray, long long P) {
  if (P <= 0) return set_error("P must be > 0");
  if (!pts) {
    if (!rays_o || !rays_d || !z) return set_error("give either pts or (rays_o, rays_d, z)");
    if (n_per_ray <= 0 || P % n_per_ray) return set_error("P must be a multiple of n_per_ray");
  }
  return 0;
}

int set_cluster_width(int v) {
  if (v != 1 && v != 2 && v != 3 && v != -2)
    return set_error("cluster must be 1, 2 (multicast), 3 (1 with the unrolled issuer loop) or -2 (cta_group::2 pairs)");
  g_cluster = v;
  return 0;
}

}  // namespace emap

using namespace emap;

extern "C" int emap_udf_forward(const emap_net_desc* net, const void* packed, int precision,
                                const float* pts, const float* rays_o, const float* rays_d,
                                const float* z, int32_t n_per_ray, int64_t P, float* udf_out,
                                float* pe_out, void* stream) {
  if (check_net(net)) return 1;
  if (!packed || !udf_out) return set_error("emap_udf_forward: NULL pointer");
  if (check_points(pts, rays_o, rays_d, z, n_per_ray, P)) return 1;
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.packed = (const uint8_t*)packed; a.pts = pts; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z;
  a.n_per_ray = n_per_ray; a.P = P; a.udf_out = udf_out; a.pe_out = pe_out;
  return dispatch<0>(net, precision, a, (cudaStream_t)stream);
}

extern "C" int emap_udf_forward_grad(const emap_net_desc* net, const void* packed, int precision,
                                     const float* pts, const float* rays_o, const float* rays_d,
                                     const float* z, int32_t n_per_ray, int64_t P, float* udf_out,
                                     float* grad_out, void* stream) {
  if (check_net(net)) return 1;
  if (!packed || !udf_out || !grad_out) return set_error("emap_udf_forward_grad: NULL pointer");
  if (check_points(pts, rays_o, rays_d, z, n_per_ray, P)) return 1;
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.packed = (const uint8_t*)packed; a.pts = pts; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z;
  a.n_per_ray = n_per_ray; a.P = P; a.udf_out = udf_out; a.grad_out = grad_out;
  return dispatch<1>(net, precision, a, (cudaStream_t)stream);
}

// K1b stage 1: dual forward (value + one tangent along d_grad) of layers 0..7 with fp16 stashes.
extern "C" int emap_bwd_dual_forward(const emap_net_desc* net, const void* packed, int precision,
                                     const float* pts, const float* rays_o, const float* rays_d,
                                     const float* z, int32_t n_per_ray, int64_t P, const float* d_grad,
                                     const float* scales, void* st_u0, void* st_u, void* stream) {
  if (check_net(net)) return 1;
  if (!packed || !st_u0 || !st_u) return set_error("emap_bwd_dual_forward: NULL pointer");
  if (check_points(pts, rays_o, rays_d, z, n_per_ray, P)) return 1;
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.packed = (const uint8_t*)packed; a.pts = pts; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z;
  a.n_per_ray = n_per_ray; a.P = P; a.gbar = d_grad; a.bwd_scales = scales;
  a.st_u0 = (__half*)st_u0; a.st_u = (__half*)st_u;
  return dispatch<2>(net, precision, a, (cudaStream_t)stream);
}

// K1b stage 1, shared-forward variant: when the training forward (emap_udf_forward_grad_rev with stash
// pointers) has already written the VALUE rows of st_u0 / st_u, only the tangent rows along d_grad remain:
// one row per point (128 points per tile), no softplus -- about half the work of the dual forward.
extern "C" int emap_bwd_tangent_forward(const emap_net_desc* net, const void* packed, const float* pts,
                                        const float* rays_o, const float* rays_d, const float* z,
                                        int32_t n_per_ray, int64_t P, const float* d_grad,
                                        const float* scales, void* st_u0, void* st_u, void* stream) {
  if (check_net(net)) return 1;
  if (!packed || !st_u0 || !st_u) return set_error("emap_bwd_tangent_forward: NULL pointer");
  if (check_points(pts, rays_o, rays_d, z, n_per_ray, P)) return 1;
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.packed = (const uint8_t*)packed; a.pts = pts; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z;
  a.n_per_ray = n_per_ray; a.P = P; a.gbar = d_grad; a.bwd_scales = scales;
  a.st_u0 = (__half*)st_u0; a.st_u = (__half*)st_u;
  a.dbg_clk = g_dbg_clk;          // (NULL unless emap_debug_set_clk_buffer: clock64 timeline of block 0)
  // fp16 packs: stash rows through shared memory by TMA ("tan_tma" 0 or "cluster" 3: the register-staged form)
  if (net->elem_type == 0 && g_tan_tma && g_cluster != 3) {
    if (make_stash_map(a.stash_map, st_u, P, 128)) return 1;
    return launch<1, 3, __half, 1>(a, (cudaStream_t)stream);
  }
  if (net->elem_type == 0) return launch<1, 3, __half, 3>(a, (cudaStream_t)stream);
  return launch<1, 3, __nv_bfloat16, 1>(a, (cudaStream_t)stream);
}

// Debug / test hook: run the forward (mode 0) or forward+grad (mode 1) kernel and additionally dump
// the de-scaled accumulators of tile 0, layer by layer: dbg_acc[9][128][256] floats.
extern "C" int emap_debug_mlp(const emap_net_desc* net, const void* packed, int precision, int mode,
                              const float* pts, int64_t P, float* udf_out, float* grad_out,
                              float* dbg_acc, void* stream) {
  if (check_net(net)) return 1;
  if (!packed || !udf_out || !pts) return set_error("emap_debug_mlp: NULL pointer");
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  a.packed = (const uint8_t*)packed; a.pts = pts; a.P = P; a.udf_out = udf_out; a.grad_out = grad_out;
  a.dbg_acc = dbg_acc;
  a.dbg_clk = g_dbg_clk;
  if (mode == 0) return dispatch<0>(net, precision, a, (cudaStream_t)stream);
  if (!grad_out) return set_error("emap_debug_mlp: grad_out required for mode 1");
  return dispatch<1>(net, precision, a, (cudaStream_t)stream);
}

extern "C" int emap_set_option(const char* name, int value) {
  if (!name) return set_error("option name is NULL");
  if (!strcmp(name, "cluster")) return set_cluster_width(value);
  if (!strcmp(name, "dbg")) { emap::g_dbg_flags = value; return 0; }
  if (!strcmp(name, "dbg_iter")) { emap::g_dbg_iter = value; return 0; }
  if (!strcmp(name, "dynamic_tiles")) { emap::g_dynamic_tiles = value; emap::rev::set_dynamic(value); return 0; }
  if (!strcmp(name, "dw_lbo")) return emap::dw::set_desc_strides(0, value);      // bring-up of mlp_dw.cu's descriptors
  if (!strcmp(name, "dw_sbo")) return emap::dw::set_desc_strides(1, value);
  if (!strcmp(name, "tan_tma")) { emap::g_tan_tma = value; return 0; }
  if (!strcmp(name, "rg_flags")) return emap::rg::set_flags(value);   // K1r experiment switches (mlp_rg.cu)
  return set_error("unknown option '%s'", name);
}

extern "C" int emap_debug_set_clk_buffer(void* dev_buf_144_int64) {
  emap::g_dbg_clk = (long long*)dev_buf_144_int64;
  return 0;
}
