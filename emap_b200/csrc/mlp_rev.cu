// K1b stage 2: fused reverse sweep of the dual network on the tensor cores.
//
// Given the stash of the dual forward (emap_bwd_dual_forward: U_{l+1} = (h_{l+1} ; hdot_{l+1}), from which
// sigma_l = 1 - exp(-100 h_{l+1}) and adot_l softplus''(a_l) = 100 hdot_{l+1} (1 - sigma_l) follow) and the
// output-layer pull-back (emap_bwd_top: alpha_8, alphadot_8 per point), one persistent kernel walks the
// layers 7 -> 0 per tile of 64 points x {alpha, alphadot} rows:
//     alpha_l    = eta_{l+1} sigma_l + etadot_{l+1} adot_l softplus''(a_l)      (value row)
//     alphadot_l = etadot_{l+1} sigma_l                                        (tangent row)
//     [eta_l ; etadot_l] = [alpha_l ; alphadot_l] W_l                           (tcgen05.mma, W_l^T images)
// and stashes A_l = [alpha_l ; alphadot_l] (fp16, row-major) for the weight-gradient GEMMs
// dW_l = A_l^T U_l.  Same skeleton as mlp_tc.cu (bulk-copy weight ring, one MMA-issuing warp, 16
// epilogue warps converting 64-column chunks in order so the next layer's MMA overlaps); single-MMA
// fp16 arithmetic with fp32 accumulation.  Replaces the autograd reverse pass through
// src/models/udf_model.py:90-135 (loss.backward(), runner_udf.py:167).
#include "common.cuh"
#include "host.h"

namespace emap {

namespace rev {

constexpr int kEpiWarps = 16;
constexpr int kProducerWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + 1;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kChunkBytes = 16384;
constexpr int kStages = 4;             // ring stages of 32 KiB: one [256 x 64] W^T operand (both N halves)
constexpr int kRingStageBytes = 2 * kStageBytes;
constexpr int kRevLayers = 7;          // MMA layers l = 7..1
constexpr int kRevParts = kRevLayers * 4;

struct Smem {
  static constexpr int a = 0;                                   // [4 chunks][128 x 64] fp16 SW128
  static constexpr int ring = a + 4 * kChunkBytes;
  static constexpr int bars = ring + kStages * kRingStageBytes;
  static constexpr int total = bars + 256 + 1024;
};

struct Args {
  const uint8_t* packed;
  const float* coef;        // [2P] alpha_8 (rows [0,P)), alphadot_8 (rows [P,2P))
  const __half* st_u;       // [8][2P,256] dual activations (h ; hdot) = inputs of layers 1..8
  __half* st_a;             // [8][2P,256]  out: A_l, l = 0..7
  long long P;
  int num_tiles, iters;
  unsigned int* tile_counter;   // dynamic tile scheduling (as in mlp_rg.cu / mlp_tc.cu); NULL = static round robin
};

__device__ __forceinline__ uint32_t pack2h(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ROLL: the issuer's layer loop rolled (instruction-cache footprint, see mlp_tc.cu / mlp_rg.cu)
template <bool ROLL>
__global__ void __launch_bounds__(kThreads, 1) mlp_rev_kernel(const Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const PackedHeader* hdr = reinterpret_cast<const PackedHeader*>(args.packed);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int out3 = (int)hdr->out_dim[kSkipLayer - 1];          // 256 - pe: valid columns of alpha_3

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* full = bars;            // [8]
  uint64_t* empty = bars + 8;       // [8]
  uint64_t* a_ready = bars + 16;    // [4]
  uint64_t* acc_full = bars + 20;   // [2]
  uint64_t* acc_empty = bars + 22;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  // tile schedule: completion k of sched_ready publishes the tile of iteration k in sched_tile[k & 1] (one thread
  // of epilogue warp 0, during stage 1 of iteration k-1); a tile index >= num_tiles ends every role's loop
  uint64_t* sched_ready = bars + 26;
  volatile int* sched_tile = reinterpret_cast<volatile int*>(bars + 27);

  if (warp == kProducerWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int c = 0; c < 4; ++c) mbar_init(&a_ready[c], kEpiWarps);
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
    mbar_init(sched_ready, 1);
    sched_tile[0] = (int)blockIdx.x;
    fence_barrier_init();
    mbar_arrive(sched_ready);             // completion 0: iteration 0 runs tile blockIdx.x
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // W_l^T images, l = 7..1, K chunk by K chunk, are contiguous from hdr->reserved[2] (pack.cu:
    // build_rev_items): one 32 KiB bulk copy per (layer, K chunk); uniform control flow, elected issue.
    const uint8_t* img = args.packed + hdr->reserved[2];
    uint8_t* ring = smem + Smem::ring;
    uint32_t stage = 0, round = 0;
    for (int iter = 0;; ++iter) {
      mbar_wait(sched_ready, (uint32_t)iter & 1, 560);
      if (sched_tile[iter & 1] >= args.num_tiles) break;
#pragma unroll 1
      for (int i = 0; i < kRevParts; ++i) {
        if (round > 0) mbar_wait(&empty[stage], (round - 1) & 1, 100 + (int)stage, i);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], kRingStageBytes);
          bulk_g2s(ring + stage * kRingStageBytes, img + (size_t)i * kRingStageBytes, kRingStageBytes, &full[stage]);
        }
        __syncwarp();
        if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
      }
    }
  } else if (warp == kMmaWarp) {
    const uint32_t a_addr = smem_u32(smem + Smem::a);
    const uint32_t ring_addr = smem_u32(smem + Smem::ring);
    const uint32_t idesc = make_idesc_f16(128, 256, 0);
    uint32_t stage = 0, round = 0;
    for (int iter = 0;; ++iter) {
      mbar_wait(sched_ready, (uint32_t)iter & 1, 561);
      if (sched_tile[iter & 1] >= args.num_tiles) break;
#pragma unroll (ROLL ? 1 : kRevLayers)
      for (int j = 0; j < kRevLayers; ++j) {
        const int buf = j & 1;
        {
          const uint32_t started = (uint32_t)iter * (buf ? 3u : 4u) + (uint32_t)(j >> 1);
          if (started > 0) mbar_wait(&acc_empty[buf], (started - 1) & 1, 200 + buf, j);
        }
        const uint32_t d = tmem_base + (uint32_t)buf * 256u;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(&a_ready[kc], ((uint32_t)iter * 7u + (uint32_t)j) & 1, 300 + kc, j);
          mbar_wait(&full[stage], round & 1, 400 + (int)stage, j * 4 + kc);
          tc_fence_after();
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr + kc * kChunkBytes);
          const uint64_t bdesc = make_sw128_kmajor_desc(ring_addr + stage * kRingStageBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kc == 0 && k == 0) ? 0u : 1u);
            umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == (uint32_t)kStages) { stage = 0; ++round; }
        }
        if (elect_one()) umma_commit(&acc_full[buf]);
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue warps ================================
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;
    const int t2 = lane & 1;                                  // 0: alpha (value) row, 1: alphadot row
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* A = smem + Smem::a;
    const float* w8 = reinterpret_cast<const float*>(args.packed + hdr->weff_layer_off[8]);
    const size_t P = (size_t)args.P;

    const bool scheduler = (warp == 0 && lane == 0);
    for (int iter = 0;; ++iter) {
      mbar_wait(sched_ready, (uint32_t)iter & 1, 562);
      const long long tile = (long long)sched_tile[iter & 1];
      if (tile >= args.num_tiles) break;
      int next_tile = 0;
      if (scheduler) {
        const long long nt = args.tile_counter ? (long long)gridDim.x + (long long)atomicAdd(args.tile_counter, 1u)
                                               : tile + (long long)gridDim.x;
        next_tile = (nt < (long long)args.num_tiles) ? (int)nt : args.num_tiles;
      }
      const long long pt = tile * 64 + q * 16 + (lane >> 1);
      const bool ok = (tile < args.num_tiles) && (pt < args.P);
      const size_t pc = ok ? (size_t)pt : 0;
      const size_t rowg = (t2 ? P : 0) + pc;
      // cotangents of a8 / adot8 for this point
      const float c_v = ok ? args.coef[pc] : 0.f;
      const float c_t = ok ? args.coef[P + pc] : 0.f;

      // stage j = -1 builds A_7 from the output-layer pull-back; stages j = 0..6 from the accumulators
#pragma unroll 1
      for (int j = -1; j < kRevLayers; ++j) {
        const int lt = 6 - j;                   // layer whose A_l = [alpha ; alphadot] this stage produces
        const int buf = j & 1;
        // the stash reads of this stage do not depend on the MMA: issue them before waiting for it
        uint32_t uw_all[4][8];
        {
          const __half* u_pre = args.st_u + (size_t)lt * 2 * P * 256 + rowg * 256 + sub * 16;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) ldg256(u_pre + c4 * 64, uw_all[c4]);
          // (an L2 prefetch of the next stage's rows from here was measured: 7.5 vs 7.15 ms -- slower, removed)
        }
        if (j >= 0) {
          mbar_wait(&acc_full[buf], ((uint32_t)iter * (buf ? 3u : 4u) + (uint32_t)(j >> 1)) & 1, 500 + buf, j);
          tc_fence_after();
          if (j == 1 && scheduler) {        // all 16 warps are past stage 0 of this tile, i.e. past its schedule wait
            sched_tile[(iter + 1) & 1] = next_tile;
            mbar_arrive(sched_ready);
          }
        }
        // own row of U_{lt+1}: h (value lane) or hdot (tangent lane); the partner's comes by shuffle
        __half* a_out = args.st_a + (size_t)lt * 2 * P * 256 + rowg * 256;
        const int ncols = (lt == kSkipLayer - 1) ? out3 : 256;
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
          const int col0 = chunk * 64 + sub * 16;
          float own[16];
          if (j >= 0) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(lane_taddr + (uint32_t)(buf * 256 + col0), r);
            tmem_wait_ld();
#pragma unroll
            for (int k = 0; k < 16; ++k) own[k] = __uint_as_float(r[k]) * kInvWeightScale;
          } else {
            const float cf = t2 ? c_t : c_v;
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(w8 + col0 + k));
              own[k] = cf * w.x; own[k + 1] = cf * w.y; own[k + 2] = cf * w.z; own[k + 3] = cf * w.w;
            }
          }
          // The lane pair (value row, tangent row) splits the 16 columns: the value lane evaluates columns
          // 0-7 of BOTH rows, the tangent lane columns 8-15, so sigma = 1 - exp(-100 h) is formed once per
          // (point, column).  12 shuffles bring the partner row's adjoints / activations for the lane's 8
          // columns, 4 more return the packed results to the row that owns them.
          uint32_t outp[8];
          const uint32_t (&uw)[8] = uw_all[chunk];
          // (selects are kept to the exchange itself: what is sent, and which of {kept, received} is the value /
          //  tangent row -- the arithmetic below is the same instruction stream for both lanes of a pair)
          float eta[8], etad[8];                                   // adjoints of the value / tangent row, my 8 columns
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float got = __shfl_xor_sync(0xffffffffu, t2 ? own[k] : own[8 + k], 1);
            const float kept = t2 ? own[8 + k] : own[k];
            eta[k] = t2 ? got : kept;
            etad[k] = t2 ? kept : got;
          }
          uint32_t hvw[4], hdw[4];                                 // h (value row) / hdot (tangent row), my 8 columns
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t got = __shfl_xor_sync(0xffffffffu, t2 ? uw[w] : uw[4 + w], 1);
            const uint32_t kept = t2 ? uw[4 + w] : uw[w];
            hvw[w] = t2 ? got : kept;
            hdw[w] = t2 ? kept : got;
          }
          uint32_t pa[4], pd[4];                                   // packed alpha / alphadot of my 8 columns
          // only layer 3 has fewer than 256 outputs (its last columns are the skip input's PE part)
          const bool partial = (ncols != 256);                     // uniform over the CTA
          const int nlive = ncols - (col0 + t2 * 8);               // live columns of my half (>= 8: all)
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float2 hv = __half22float2(*reinterpret_cast<const __half2*>(&hvw[w]));
            const float2 hd = __half22float2(*reinterpret_cast<const __half2*>(&hdw[w]));
            const float hve[2] = {hv.x, hv.y}, hde[2] = {hd.x, hd.y};
            float al[2], ad[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int k = 2 * w + e;                             // column inside my half
              // sigma = softplus'(a) = 1 - exp(-100 h);  adot * softplus''(a) = 100 * hdot * (1 - sigma)
              const float one_m_s = __expf(-kSoftplusBeta * hve[e]);
              const float sg = 1.0f - one_m_s;
              al[e] = fmaf(etad[k], kSoftplusBeta * hde[e] * one_m_s, eta[k] * sg);       // alpha
              ad[e] = etad[k] * sg;                                                       // alphadot
              if (partial && k >= nlive) { al[e] = 0.f; ad[e] = 0.f; }
            }
            pa[w] = pack2h(al[0], al[1]);
            pd[w] = pack2h(ad[0], ad[1]);
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const uint32_t got = __shfl_xor_sync(0xffffffffu, t2 ? pa[w] : pd[w], 1);   // the partner row's share
            outp[w] = t2 ? got : pa[w];                            // columns 0-7 of my row
            outp[4 + w] = t2 ? pd[w] : got;                        // columns 8-15
          }
          if (lt >= 1) {
#pragma unroll
            for (int g = 0; g < 2; ++g)
              *reinterpret_cast<uint4*>(A + chunk * kChunkBytes + (uint32_t)row * 128u +
                                        (uint32_t)((((sub * 2 + g) ^ (row & 7)) & 7) << 4)) =
                  make_uint4(outp[g * 4], outp[g * 4 + 1], outp[g * 4 + 2], outp[g * 4 + 3]);
          }
          if (lt >= 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[chunk]);
          }
          // the stash row goes out AFTER the hand-off: a global store in front of the fence (MEMBAR.ALL.CTA +
          // FENCE.VIEW.ASYNC) makes every hand-off wait for an L2 round trip
          if (ok) stg256(a_out + col0, outp);
        }
        if (j >= 0) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

static int g_dynamic = 1;
int set_dynamic(int v) { g_dynamic = v; return 0; }
// emap_set_option("rev_rolled", 0): the unrolled issuer loop of round 1 (A/B switch; 7.2 vs 7.1 ms rolled)
static int g_rolled = 1;
int set_rolled(int v) { g_rolled = v; return 0; }

}  // namespace rev
}  // namespace emap

using namespace emap;

extern "C" int emap_bwd_reverse_sweep(const emap_net_desc* net, const void* packed, const float* coef,
                                      const void* st_u, void* st_a, int64_t P, void* stream) {
  if (check_net(net)) return 1;
  if (net->elem_type != 0) return set_error("emap_bwd_reverse_sweep: fp16 operand images required");
  if (!packed || !coef || !st_u || !st_a || P <= 0) return set_error("emap_bwd_reverse_sweep: bad arguments");
  rev::Args a;
  a.packed = (const uint8_t*)packed; a.coef = coef; a.st_u = (const __half*)st_u;
  a.st_a = (__half*)st_a; a.P = P;
  const long long tiles = (P + 63) / 64;
  a.num_tiles = (int)tiles;
  int grid = sm_count();
  if (tiles < grid) grid = (int)tiles;
  a.iters = (int)((tiles + grid - 1) / grid);
  a.tile_counter = rev::g_dynamic ? tile_counter((cudaStream_t)stream) : nullptr;
  static bool attr_done = false;
  if (!attr_done) {
    EMAP_CUDA(cudaFuncSetAttribute(rev::mlp_rev_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rev::Smem::total));
    EMAP_CUDA(cudaFuncSetAttribute(rev::mlp_rev_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rev::Smem::total));
    attr_done = true;
  }
  if (rev::g_rolled & 1) rev::mlp_rev_kernel<true><<<grid, rev::kThreads, rev::Smem::total, (cudaStream_t)stream>>>(a);
  else rev::mlp_rev_kernel<false><<<grid, rev::kThreads, rev::Smem::total, (cudaStream_t)stream>>>(a);
  EMAP_CUDA(cudaGetLastError());
  return 0;
}
