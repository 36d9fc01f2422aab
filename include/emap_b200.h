/* emap_b200 -- C ABI of the B200-native EMAP volume-rendering hot path.
 *
 * Plain C, raw device pointers + explicit sizes, a cudaStream_t (as void*) last, int status
 * return (0 = ok, non-zero = error; text via emap_last_error()).  No torch types cross this
 * boundary.  All pointers are DEVICE pointers unless the name ends in _host.
 *
 * The reference (cvg/EMAP) has no FFI layer: its "interface" for this path is the Python methods
 * of src/models/udf_model.py and src/models/udf_renderer_blending.py.  Each entry point below
 * names the reference method(s) it replaces (file:line relative to the reference root); the Python
 * classes in emap_b200/ (same names / kwargs / return dicts as the reference) are thin shims over
 * these calls.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 */
#ifndef EMAP_B200_H
#define EMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMAP_ABI_VERSION 2

/* Network description == the `model.udf_network` conf block (confs/ABC.conf:65-77) restricted
 * to the topology the kernels are built for: d_in=3, d_hidden=256, n_layers=8, skip_in=[4],
 * d_out=1, weight_norm=True.  Anything else is rejected with an error (never silently emulated). */
typedef struct emap_net_desc {
  int32_t multires;   /* number of PE frequencies L, 0..10  (udf_model.py:30-33)            */
  int32_t udf_type;   /* 0 = "abs", 1 = "square", 2 = "sdf" (udf_model.py:82-88)           */
  float   scale;      /* UDFNetwork.scale                   (udf_model.py:36,91,108)       */
  int32_t elem_type;  /* tensor-core operand type: 0 = fp16, 1 = bf16                       */
} emap_net_desc;

/* MLP arithmetic mode */
#define EMAP_PREC_FP32X3 3 /* fp32-faithful: 3 split-fp16 tcgen05 MMAs (hi*hi + lo*hi + hi*lo), fp32 accum */
#define EMAP_PREC_HALF   1 /* one fp16 (or bf16) tcgen05 MMA, fp32 accumulate                              */

/* Device-side status word (an int32 in device memory owned by the caller, zeroed by the caller): kernels OR
 * bits into it instead of the reference's `pdb.set_trace()` NaN guards; the host shim polls it once per step
 * and raises (udf_renderer_blending.py:102-107, :346-351, :632-633).                                        */
#define EMAP_STATUS_NAN_SAMPLES     1 /* NaN among the new z samples of an up-sampling step (:102, :346)      */
#define EMAP_STATUS_NAN_EIKONAL     2 /* NaN gradient_error in render_core (:632)                             */
#define EMAP_STATUS_NONFINITE_GRAD  4 /* non-finite parameter gradient out of the backward                    */

/* Flat parameter buffer layout (fp32), identical to list(UDFNetwork.parameters()) order:
 *   for l in 0..8: bias[out_l], g[out_l] ("original0"), v[out_l*in_l] ("original1")
 *   in  = [pe,256,256,256,256,256,256,256,256], out = [256,256,256,256-pe,256,256,256,256,1],
 *   pe = 3+6*multires.   (udf_model.py:39-76; checkpoint keys SURVEY §5)                          */
size_t emap_flat_param_count(const emap_net_desc* net);

const char* emap_last_error(void);
int emap_abi_version(void);

/* ---- K0: weight-norm fold + tensor-core operand packing ------------------------------------
 * replaces: torch weight_norm recomputation on every forward (udf_model.py:74; 63 calls/render).
 * `packed` must hold emap_packed_size(net) bytes; it is rewritten whenever parameters change.   */
size_t emap_packed_size(const emap_net_desc* net);
int emap_wn_fold(const emap_net_desc* net, const float* flat_params, void* packed, void* stream);

/* ---- K1: fused PE + 9-layer MLP forward -------------------------------------------------------
 * replaces: Embedder.embed (embedder.py:34) + UDFNetwork.forward/.udf (udf_model.py:90-116).
 * Points are given either explicitly (pts[P,3]) or implicitly as rays: P = n_rays*n_per_ray,
 * point i = rays_o[i/n] + rays_d[i/n] * z[i]  (udf_renderer_blending.py:812, :360, :448).
 * udf_out[P]; pe_out (optional) [P, 3+6L] in the reference column order.                         */
int emap_udf_forward(const emap_net_desc* net, const void* packed, int precision,
                     const float* pts, const float* rays_o, const float* rays_d, const float* z,
                     int32_t n_per_ray, int64_t P, float* udf_out, float* pe_out, void* stream);

/* ---- K1g: forward + d udf / d x in one pass (forward-mode tangents on the tensor cores) -------
 * replaces: UDFNetwork.forward + UDFNetwork.gradient (udf_model.py:121-135), i.e. the second
 * forward and the autograd.grad call at udf_renderer_blending.py:457-461.  grad_out[P,3].        */
int emap_udf_forward_grad(const emap_net_desc* net, const void* packed, int precision,
                          const float* pts, const float* rays_o, const float* rays_d,
                          const float* z, int32_t n_per_ray, int64_t P, float* udf_out,
                          float* grad_out, void* stream);

/* ---- K1r: the same result as K1g by REVERSE mode (value-only forward that keeps softplus' of every
 * layer, then the adjoint sweep with the W^T operand images): 6 F executed per point instead of 12 F.
 * replaces: the same reference lines as emap_udf_forward_grad.  `scratch` = emap_rgrad_scratch_bytes()
 * bytes of device memory (448 KiB per SM, rewritten tile after tile -> L2-resident), reusable across
 * calls on one stream.  Selected by the host shim with EMAP_GRAD_MODE=reverse / ops.set_grad_mode().  */
size_t emap_rgrad_scratch_bytes(void);
/* st_u0 / st_u (both or neither; training only): the backward's stashes (see emap_bwd_dual_forward) --
 * the forward then writes their VALUE rows [0,P) (PE and h_1..h_8 as fp16) and the backward only has to
 * add the tangent rows with emap_bwd_tangent_forward instead of re-running the dual forward.        */
int emap_udf_forward_grad_rev(const emap_net_desc* net, const void* packed, int precision,
                              const float* pts, const float* rays_o, const float* rays_d,
                              const float* z, int32_t n_per_ray, int64_t P, float* udf_out,
                              float* grad_out, void* scratch, size_t scratch_bytes, void* st_u0,
                              void* st_u, void* stream);

/* ---- K1b: backward of (udf, d udf/dx) w.r.t. the 462,980 MLP parameters -------------------------
 * replaces: autograd through UDFNetwork.forward + .gradient(create_graph=True)
 * (udf_model.py:90-135; loss.backward() at runner_udf.py:167).  All *_half pointers are fp16 device buffers;
 * "dual" tensors hold value rows [0,P) and tangent rows [P,2P).
 *
 * Loss scaling: the stashes and MMA operands of the backward are fp16, raw loss cotangents are not in its
 * range (|dL/dgrad| ~ 1e-8 at production batch sizes).  emap_bwd_cotangent_scales derives two powers of two
 * from the maxima of the cotangents on the device (no host sync): scales[0] = S_g (the tangent direction is
 * S_g d_grad, max in [0.5,1)), scales[1] = S_u (what the sweep accumulates is S_u dL/dtheta; max(S_u |d_udf|,
 * S_u |d_grad|) in [1/8,1/4)), scales[2] = S_u/S_g, scales[3] = 1/S_u; scales[4..7] scratch.  The stages
 * below take that buffer (NULL = all ones); emap_bwd_finish removes S_u.                                  */
int emap_bwd_cotangent_scales(const float* d_udf /*[P] or NULL*/, const float* d_grad /*[P,3] or NULL*/,
                              int64_t P, float* scales8, void* stream);
/* Workspace of the weight-gradient stage: per-CTA partial dW / db of emap_bwd_weight_grads and the per-block
 * partial dW_8 / db_8 of emap_bwd_top (~290 MB, independent of P; reusable across calls on one stream).      */
size_t emap_bwd_workspace_bytes(void);
/* output-layer pull-back: coef[p] = S_u d_udf f'(a8)/scale + (S_u/S_g) f''(a8) adot8, coef[P+p] = (S_u/S_g) f'(a8);
 * also the output layer's own gradient dW_8 = sum_p coef[p] U8[p] + coef[P+p] U8[P+p], db_8 = sum_p coef[p] as
 * per-block partials in the workspace (replaces a library GEMV + two reductions).                          */
int emap_bwd_top(const emap_net_desc* net, const void* U8_half, const float* w8, const float* b8,
                 const float* d_udf /*[P] or NULL*/, const float* scales, int64_t P, float* coef /*[2P]*/,
                 void* workspace, void* stream);
/* Fused tensor-core stages of K1b (hand-written tcgen05 kernels on the K1 skeleton):
 *  emap_bwd_dual_forward : layers 0..7 of the dual network (value + ONE tangent along d_grad) with
 *     fp16 stashes  st_u0[2P,64] (dual PE, kernel column order), st_u[8][2P,256] (inputs of layers 1..8:
 *     h_{l+1} in rows [0,P), hdot_{l+1} in rows [P,2P)).  No sigma / adot stash is needed:
 *     softplus'(a_l) = 1 - exp(-100 h_{l+1}) and adot_l softplus''(a_l) = 100 hdot_{l+1} (1 - sigma_l).
 *  emap_bwd_reverse_sweep: layers 7..0 of the reverse sweep from coef[2P] (emap_bwd_top) and st_u;
 *     writes st_a[8][2P,256] = [alpha_l ; alphadot_l].  The weight gradients are then the contractions
 *     dW_l = A_l^T U_l of emap_bwd_weight_grads, finished by emap_bwd_finish.
 *  The stashes move through shared memory with the TMA engine (reverse sweep: both; tangent forward: st_u;
 *  training forward emap_udf_forward_grad_rev: the value rows of st_u; weight gradients: all three): st_u, st_a
 *  and st_u0 must be 16-byte aligned, dense in the layouts above (any torch allocation is).  Ragged sizes
 *  (P not a multiple of the 64- / 128-point tiles) are handled by the tensor maps (zero fill / clipping).      */
int emap_bwd_dual_forward(const emap_net_desc* net, const void* packed, int precision,
                          const float* pts, const float* rays_o, const float* rays_d, const float* z,
                          int32_t n_per_ray, int64_t P, const float* d_grad, const float* scales,
                          void* st_u0, void* st_u, void* stream);
int emap_bwd_reverse_sweep(const emap_net_desc* net, const void* packed, const float* coef,
                           const void* st_u, void* st_a, int64_t P, void* stream);
/* Shared-forward variant of stage 1: the value rows of st_u0 / st_u were written by the training forward
 * (emap_udf_forward_grad_rev with stash pointers); this adds the tangent rows [P,2P) along d_grad (NULL =
 * zero tangent): one row per point, sigma_l recovered from the stashed h_{l+1}, single fp16 MMA.       */
int emap_bwd_tangent_forward(const emap_net_desc* net, const void* packed, const float* pts,
                             const float* rays_o, const float* rays_d, const float* z, int32_t n_per_ray,
                             int64_t P, const float* d_grad, const float* scales, void* st_u0, void* st_u,
                             void* stream);
/* The weight-gradient contractions dW_l = A_l^T U_l, l = 0..7, and the bias sums db_l = sum_{p<P} A_l[p,:], on the
 * tensor cores (mlp_dw.cu: TMA-staged row tiles as MN-major tcgen05 operands, one [256 x 256] fp32 accumulator
 * per CTA in TMEM): st_a [8][2P,256], st_u0 [2P,64], st_u [8][2P,256] fp16 -> per-CTA partials in the workspace.
 * replaces the `mm` chain of autograd's addmm backward (udf_model.py:102 under loss.backward()).
 * Returns the number of partials (> 0) for emap_bwd_finish, or -1 with emap_last_error() set.                */
int emap_bwd_weight_grads(const emap_net_desc* net, const void* st_a, const void* st_u0, const void* st_u,
                          int64_t P, void* workspace, size_t workspace_bytes, void* stream);
/* Final stage: fixed-order sum of the partials (deterministic, no atomics), the kernels' PE column order undone,
 * weight-norm backward, scatter into the flat gradient (same layout as the flat parameters), loss scale S_u
 * removed (scales[3]; NULL = 1).  status (optional, device int32): bit EMAP_STATUS_NONFINITE_GRAD is set when a
 * parameter gradient is not finite (e.g. an fp16 overflow of the scaled sweep).                              */
int emap_bwd_finish(const emap_net_desc* net, const float* flat_params, const void* workspace, int32_t n_parts,
                    const float* scales, float* flat_grad, int32_t* status, void* stream);
/* byte offsets inside the packed buffer: out[0] = 100*bias table, out[1..9] = W_eff of layer 0..8. */
int emap_packed_offsets(const emap_net_desc* net, uint32_t* out10);

/* ---- per-ray kernels (one warp per ray; <= 512 samples per ray, <= 64 new samples per step) ----
 * z = near + (far-near)*lin + t_rand*2/n                   (udf_renderer_blending.py:705-720)
 * near/far: device scalars [1] (near_is_per_ray=0) or [B] (=1); lin = torch.linspace(0,1,n) on
 * the device; t_rand [B] = rand-0.5 (NULL when perturb == 0).                                    */
int emap_coarse_z(const float* near, const float* far, int32_t near_is_per_ray, const float* lin,
                  const float* t_rand, int32_t B, int32_t n, float* z_out, void* stream);

/* One hierarchical up-sampling step, fused:
 *   (1) if ka > 0: merge the pending samples (z_add, udf_add)[B,ka] into (z_in, udf_in)[B,n]
 *       -> (z_out, udf_out)[B,n+ka]                      replaces cat_z_vals  (:355-377)
 *   (2) if k > 0: density -> transmittance -> weights -> inverse CDF at the fixed quantiles u[k]
 *       -> z_new[B,k] (ascending), inds_out[B,k] (searchsorted result, int64, optional),
 *       weights_out[B,n+ka-1] (optional)   replaces up_sample_unbias (:228-353) [mode 0],
 *       up_sample_no_occ_aware (:920-975) [mode 1] and sample_pdf(det=True) (:69-109);
 *       mode 2 = sample_pdf alone: udf_in then holds the weights [B,n-1] as given.
 * alpha_type: 0 "numerical", 1 "theorical" (:399-414).  sample_dist: device scalar.  gamma_dev
 * (optional device scalar) overrides `gamma` -- importance_sample_mix passes the learnable
 * BetaNetwork gamma (:873,:885) without a host sync.                                             */
int emap_upsample_step(const float* rays_o, const float* rays_d, const float* z_in,
                       const float* udf_in, int32_t n, const float* z_add, const float* udf_add,
                       int32_t ka, float* z_out, float* udf_out, const float* u, int32_t k,
                       float* z_new, int64_t* inds_out, float* weights_out, const float* sample_dist,
                       int32_t B, float inv_s, float beta, float gamma, const float* gamma_dev,
                       int32_t mode, int32_t alpha_type, int32_t* status /* optional, EMAP_STATUS_* */,
                       void* stream);

/* render_core before the MLP: dists, mid_z_vals          (udf_renderer_blending.py:435-446)     */
int emap_render_prep(const float* z, const float* sample_dist, int32_t B, int32_t n, float* dists,
                     float* mid_z, void* stream);

/* render_core after the MLP (:463-650): occlusion-aware alpha, transmittance scan, compositing,
 * eikonal / sparsity reductions.  scalars = device [inv_s, beta, gamma] (already exp'ed + clipped,
 * :466-472).  cos_anneal_ratio < 0 means None.  Outputs: per sample weights, alpha (optional),
 * grad_flip[B,n,3], inside_sphere, grad_mag; per ray edge, depth (un-scaled), normals[B,3];
 * partials[B,5] (fp64 scratch); reduced[5] = {gradient_error, gradient_error_near_surface,
 * sparse_error, sum(relax_inside_sphere), sum(near_surface)}.                                    */
int emap_render_core_fwd(const float* rays_o, const float* rays_d, const float* mid_z,
                         const float* dists, const float* udf, const float* grad,
                         const float* scalars, int32_t B, int32_t n, float cos_anneal_ratio,
                         float flip_saturation, float near_surface, float sparse_scale,
                         int32_t use_unbias, int32_t use_norm_grad, int32_t alpha_type,
                         float* weights, float* alpha, float* grad_flip, float* inside_sphere,
                         float* grad_mag, float* edge, float* depth, float* normals,
                         double* partials, float* reduced, int32_t* status /* optional */, void* stream);

/* backward of emap_render_core_fwd: cotangents (any may be NULL = zero) of weights[B,n], edge[B],
 * depth[B], normals[B,3] and of the three scalar reductions (device scalars) -> d_udf[B*n],
 * d_grad[B*n,3], d_scalars[3] = d/d(inv_s, beta, gamma).  `reduced` is the forward's output (its
 * mask sums are the denominators of the eikonal terms).  partials: fp64 scratch [B,3].
 * replaces: autograd through udf_renderer_blending.py:463-650.                                    */
int emap_render_core_bwd(const float* rays_o, const float* rays_d, const float* mid_z,
                         const float* dists, const float* udf, const float* grad,
                         const float* scalars, const float* reduced, int32_t B, int32_t n,
                         float cos_anneal_ratio, float flip_saturation, float near_surface,
                         float sparse_scale, int32_t use_unbias, int32_t use_norm_grad,
                         int32_t alpha_type, const float* d_weights, const float* d_edge,
                         const float* d_depth, const float* d_normals, const float* d_gerr,
                         const float* d_gerr_ns, const float* d_sparse, float* d_udf, float* d_grad,
                         double* partials, float* d_scalars, void* stream);

/* ---- callers either side of the render path ------------------------------ */
/* Line direction of edge extraction: grad [M,S,3] (S normalised UDF gradients sampled around each of
 * M near-surface voxels) -> out [M,3] = unit right-singular vector of the smallest singular value of
 * each [S,3] block (sign arbitrary, as with LAPACK).  replaces torch.linalg.svd + vh[:, -1, :] +
 * F.normalize in src/edge_extraction/extract_pointcloud.py:75-89 and :176-187.                    */
int emap_null_direction(const float* grad, int64_t M, int32_t S, float* out, void* stream);
/* Deterministic half of the ray sampler: pixels (x,y) int64 [B] (device) of ONE image -> ndc uv
 * [B,2], edge [B,1] gathered from edge_img [H,W] (device), p_cam = K^-1 (x,y,1) [B,3],
 * depth_scale = normalised p_cam.z [B,1], rays_v = R p_cam/|p_cam| [B,3], rays_o = camera centre
 * [B,3].  intr_inv3x3 (9 floats, row-major) and pose4x4 (16 floats) are HOST pointers.
 * replaces src/dataset/dataset.py:268-290 (gen_random_rays_patches_at, after the pixel draw).      */
int emap_rays_from_pixels(const int64_t* pixels_x, const int64_t* pixels_y, const float* edge_img,
                          int32_t H, int32_t W, const float* intr_inv3x3, const float* pose4x4,
                          int32_t B, float* rays_o, float* rays_v, float* edge, float* ndc_uv,
                          float* p_cam, float* depth_scale, void* stream);

/* ---- RenderingNetwork.forward as a standalone operator (src/models/udf_model.py:177-209).
 * Defined and configured in the reference but never called (render_core uses a constant edge of ones,
 * udf_renderer_blending.py:561): NOT part of render().  Fused fp32 forward: input assembly for mode 0 "idr" /
 * 1 "no_view_dir" / 2 "no_normal" incl. the view-direction encoding, n_layers Linear (+ReLU), sigmoid if
 * squeeze_out.  wt[l] = W_l^T [dims[l], dims[l+1]] row-major with weight-norm folded, bias[l] [dims[l+1]]
 * (host arrays of device pointers); out [P, d_out].  Forward only.                                          */
int emap_rendering_network_forward(const float* const* wt, const float* const* bias, const int32_t* dims,
                                   int32_t n_layers, int32_t mode, int32_t multires_view, int32_t d_feature,
                                   int32_t d_out, int32_t squeeze_out, const float* points, const float* normals,
                                   const float* view_dirs, const float* feats, int64_t P, float* out,
                                   void* stream);

/* ---- misc ------------------------------------------------------------------------------------ */
/* options -- A/B switches; every default is the measured-best setting, every alternative is bit-identical to it
 * (tests/test_gpu_rgrad.py, tests/test_gpu_mlp.py, tests/test_gpu_dw.py):
 *   "cluster" = 1|2|3|-2 : weight-stream organisation of the K1 / K1g / dual / tangent kernels.  1 (default) = every
 *               CTA streams its own copy, MMA-issuer loop ROLLED (instruction-cache footprint: K1 2.70 vs 3.04 ms);
 *               3 = the same with the unrolled issuer of round 1; 2 = multicast pairs; -2 = cta_group::2 pairs;
 *   "rg_flags" : K1r switches, default 28 = 8 (tiles handed out by a global atomic counter: 5.24 vs 6.05 ms per
 *               1 M points) | 4 (rolled MMA-issuer loop: 6.35 vs 7.26 ms in training mode) | 16 (rolled chunk loop of
 *               the reverse epilogue).  bit 0: N-split of each step's last K chunk; bit 1: persisting-L2 window on
 *               the sigma scratch (both measured, no gain); bit 5 (32): training stash stored from registers
 *               instead of by TMA from the A tile (6.0 vs 7.1 ms);
 *   "tan_tma"  : 1 (default) = the tangent forward moves its stash rows through shared memory with the TMA engine
 *               (2.26 vs 2.63 ms); 0 = register-staged;
 *   "dynamic_tiles" : 1 (default) = K1 / tangent / dual forward and reverse sweep hand tiles out dynamically;
 *   "dbg"      : timing experiments of mlp_tc.cu (0 in production);
 *   "dbg_iter" : tile iteration of block 0 that the clock64 timelines of the debug entry points stamp;
 *   "dw_lbo" / "dw_sbo" : byte strides of mlp_dw.cu's MN-major operand descriptors (8192 / 1024).             */
int emap_set_option(const char* name, int value);
/* test hook: MLP forward (mode 0) / forward+grad (mode 1) that also dumps the de-scaled
 * accumulators of tile 0, dbg_acc[9][128][256].                                                  */
/* test hook: device buffer of 144 int64 receiving clock64 stamps from emap_debug_mlp runs.        */
int emap_debug_set_clk_buffer(void* dev_buf_144_int64);
int emap_debug_mlp(const emap_net_desc* net, const void* packed, int precision, int mode,
                   const float* pts, int64_t P, float* udf_out, float* grad_out, float* dbg_acc,
                   void* stream);
/* test hook: emap_udf_forward_grad_rev that also dumps the accumulators of tile 0 after each of its 16
 * MMA steps, dbg_acc[16][128][256] (see mlp_rg.cu).                                                   */
int emap_debug_rgrad(const emap_net_desc* net, const void* packed, int precision, const float* pts,
                     const float* rays_o, const float* rays_d, const float* z, int32_t n_per_ray,
                     int64_t P, float* udf_out, float* grad_out, void* scratch, size_t scratch_bytes,
                     float* dbg_acc, void* stream);
/* test hooks on HOST memory (no GPU needed): the un-split fp32 value image `b` (0..63) of the K1r
 * reverse stream from the HOST W_eff matrix of its layer -> out_host [rows x 64] (rows returned: 256 or
 * 64; -1 on error), layer_kc_part[3] (optional) = {layer, K chunk, hi/lo part}; K1r's PE-adjoint
 * contraction itself (J_gamma^T over 16 slots, the device function compiled for the host); and the two PE
 * column maps of the kernels (kernel column / K1r slot -> reference PE index, -1 = padding).        */
int emap_debug_rg_image(const emap_net_desc* net, int b, const float* W_host, float* out_host,
                        int32_t* layer_kc_part);
int emap_debug_pe_adjoint(const float* adj16_host, int kbase, const float* x3_host, int multires,
                          float* g3_host /* accumulated into */);
int emap_debug_pe_col_to_ref(int col, int multires);
int emap_debug_rg_pe_ref(int k, int multires);

#ifdef __cplusplus
}
#endif
#endif /* EMAP_B200_H */
