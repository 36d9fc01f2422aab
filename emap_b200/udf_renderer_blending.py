"""Drop-in replacement for the reference's ``src/models/udf_renderer_blending.py``.

``UDFRendererBlending`` keeps the reference constructor kwargs (= the ``model.udf_renderer`` conf
block), the public attributes ``n_samples`` / ``n_importance`` and the ``render()`` signature and
18-key return dict (consumers ``runner_udf.py:96-162,315-408``).  Every stage runs in a
CUDA kernel behind the C ABI:

    coarse z            emap_coarse_z                (udf_renderer_blending.py:705-720)
    importance_sample   emap_udf_forward + emap_upsample_step per step   (:802-841, :228-377)
    render_core         emap_render_prep -> emap_udf_forward_grad -> emap_render_core_fwd (:418-677)

The reference's 12 host syncs per call (NaN checks that drop into pdb, ``.item()``) are gone: nothing
in ``render()`` synchronises with the device.  Its NaN guards (:102-107, :346-351, :632-633) became a device
status word the kernels OR into: ``render()`` polls it without blocking (a NaN raises ``FloatingPointError``
at the next call at the latest), ``check_numerics()`` is the blocking check for callers that synchronise
anyway (the reference runner does, once per iteration: ``runner_udf.py:164``).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .autograd import udf_forward_fn
from .render_autograd import render_core_fn


def sample_pdf(bins, weights, n_samples, det=False):
    """Inverse-CDF resampling of one step in isolation (reference: module-level ``sample_pdf``,
    udf_renderer_blending.py:69-109).  Only the deterministic branch is used by the renderer."""
    if not det:
        raise NotImplementedError("emap_b200.sample_pdf: only det=True is on the hot path")
    return ops.sample_pdf_det(bins, weights, n_samples)[0]


class UDFRendererBlending:
    def __init__(self, nerf, udf_network, deviation_network, beta_network, n_samples, n_importance,
                 n_outside, up_sample_steps, perturb, sdf2alpha_type="numerical",
                 upsampling_type="classical", sparse_scale_factor=25000, use_norm_grad_for_cosine=False,
                 use_unbias_render=True, near_surface=0.05, device="cuda"):
        self.nerf = nerf
        self.udf_network = udf_network
        self.deviation_network = deviation_network
        self.beta_network = beta_network
        self.n_samples = n_samples
        self.n_importance = n_importance
        self.n_outside = n_outside
        self.perturb = perturb
        self.up_sample_steps = up_sample_steps
        self.use_unbias_render = use_unbias_render
        self.sdf2alpha_type = sdf2alpha_type
        self.upsampling_type = upsampling_type
        self.sparse_scale_factor = sparse_scale_factor
        self.use_norm_grad_for_cosine = use_norm_grad_for_cosine
        self.near_surface = near_surface
        self.device = torch.device(device)
        if n_outside > 0 or nerf is not None:
            # reference: nerf_outside is None and n_outside = 0 in every conf (runner_base.py:88,
            # confs/*.conf); the background branch of render_core is unreachable.
            raise NotImplementedError("n_outside > 0 / NeRF background is dead code in the reference")
        if sdf2alpha_type not in ("numerical", "theorical"):
            raise ValueError(f"unknown sdf2alpha_type {sdf2alpha_type!r}")
        if upsampling_type not in ("classical", "mix"):
            raise ValueError(f"unknown upsampling_type {upsampling_type!r}")
        self._alpha_type = 0 if sdf2alpha_type == "numerical" else 1
        self._const_cache: Dict = {}
        # multi-GPU: make the two eikonal means those of the whole (sharded) batch (parallel.py)
        self.global_batch_stats = False
        # False: the stratified offsets come from the global CPU generator exactly like the reference (:719: same
        # seed -> same samples).  True: drawn on the device -- same distribution, another stream, no host->device
        # copy; required inside a CUDA-graph capture (graph.GraphedStep), where a host draw would be frozen.
        self.perturb_on_device = False

    def check_numerics(self):
        """Blocking check of the device status word (one 4-byte read); raises FloatingPointError where the
        reference would have dropped into pdb, and after a non-finite parameter gradient."""
        ops.check_status(self.device)

    # ------------------------------------------------------------------ cached device constants
    def _const(self, key, make):
        t = self._const_cache.get(key)
        if t is None:
            t = make().to(self.device)
            self._const_cache[key] = t
        return t

    def _linspace01(self, n):
        # torch.linspace evaluated on the CPU exactly as the reference does (:705), then cached
        return self._const(("lin", n), lambda: torch.linspace(0.0, 1.0, n))

    def _quantiles(self, k):
        return self._const(("u", k), lambda: torch.linspace(0.0 + 0.5 / k, 1.0 - 0.5 / k, steps=k))

    # ------------------------------------------------------------------ a9 importance sampling
    @torch.no_grad()
    def importance_sample(self, rays_o, rays_d, z_vals, sample_dist):
        """z[B,n0] -> z[B,n0+S*k]   (udf_renderer_blending.py:802-841)."""
        net = self.udf_network
        S = self.up_sample_steps
        k = self.n_importance // S
        mode = 0 if self.use_unbias_render else 1
        if not torch.is_tensor(sample_dist):
            sample_dist = torch.tensor([sample_dist], dtype=torch.float32, device=self.device)
        u = self._quantiles(k)
        cur_z = z_vals
        cur_udf, _ = udf_forward_fn(net, rays_o=rays_o, rays_d=rays_d, z=cur_z)
        cur_udf = cur_udf.view_as(cur_z)
        pend_z = pend_udf = None
        for i in range(S):
            inv_s = 64.0 * 2 ** i
            beta = 64.0 * 2 ** (i + 1)
            gamma = float(np.clip(20 * 2 ** (S - i), 20, 320))
            cur_z, cur_udf, z_new, _, _ = ops.upsample_step(
                rays_o, rays_d, cur_z, cur_udf, pend_z, pend_udf, u, k, sample_dist, inv_s, beta, gamma,
                mode, self._alpha_type)
            pend_z = z_new
            if i + 1 < S:
                pend_udf, _ = udf_forward_fn(net, rays_o=rays_o, rays_d=rays_d, z=z_new)
                pend_udf = pend_udf.view_as(z_new)
            else:
                pend_udf = None
        cur_z, _, _, _, _ = ops.upsample_step(rays_o, rays_d, cur_z, None, pend_z, None, None, 0,
                                              sample_dist, 0.0, 0.0, 0.0, mode, self._alpha_type)
        return cur_z

    @torch.no_grad()
    def importance_sample_mix(self, rays_o, rays_d, z_vals, sample_dist):
        """upsampling_type="mix" (udf_renderer_blending.py:843-918): S occlusion-unaware steps using
        the learnable BetaNetwork gamma, then one occlusion-aware step; k = n_importance // (S+1)."""
        net = self.udf_network
        S = self.up_sample_steps
        k = self.n_importance // (S + 1)
        if not torch.is_tensor(sample_dist):
            sample_dist = torch.tensor([sample_dist], dtype=torch.float32, device=self.device)
        u = self._quantiles(k)
        gamma_dev = self.beta_network.get_gamma().clip(1e-6, 1e6).detach().reshape(1).contiguous()
        cur_z = z_vals
        cur_udf, _ = udf_forward_fn(net, rays_o=rays_o, rays_d=rays_d, z=cur_z)
        cur_udf = cur_udf.view_as(cur_z)
        pend_z = pend_udf = None
        schedule = [(1, 64.0 * 2 ** i, 64.0 * 2 ** (i + 1), 0.0, gamma_dev) for i in range(S)]
        i = S - 1
        schedule.append((0, 64.0 * 2 ** i, 64.0 * 2 ** (i + 1), 20.0 if i < 4 else 10.0, None))
        for j, (mode, inv_s, beta, gamma, gdev) in enumerate(schedule):
            cur_z, cur_udf, z_new, _, _ = ops.upsample_step(
                rays_o, rays_d, cur_z, cur_udf, pend_z, pend_udf, u, k, sample_dist, inv_s, beta, gamma,
                mode, self._alpha_type, gamma_dev=gdev)
            pend_z = z_new
            if j + 1 < len(schedule):
                pend_udf, _ = udf_forward_fn(net, rays_o=rays_o, rays_d=rays_d, z=z_new)
                pend_udf = pend_udf.view_as(z_new)
            else:
                pend_udf = None
        cur_z, _, _, _, _ = ops.upsample_step(rays_o, rays_d, cur_z, None, pend_z, None, None, 0,
                                              sample_dist, 0.0, 0.0, 0.0, 0, self._alpha_type)
        return cur_z

    def _scalars(self, deviation_network, beta_network):
        """[inv_s, beta, gamma] as one differentiable 3-vector.  The drop-in scalar modules take a vectorised
        path (cat, mul, exp, clamp: the same values and the same clip masks in 4 launches instead of ~13, and
        as few again in the backward); any other module is called through the reference's own expressions."""
        from .udf_model import BetaNetwork, SingleVarianceNetwork
        if type(deviation_network) is SingleVarianceNetwork and type(beta_network) is BetaNetwork:
            key = ("clip", float(beta_network.beta_min))
            lo_hi = self._const_cache.get(key)
            if lo_hi is None:
                hi_b = min(1.0 / beta_network.beta_min, 1e6)
                lo_hi = (torch.tensor([1e-6, 1e-6, 1e-6], device=self.device),
                         torch.tensor([1e6, hi_b, 1e6], device=self.device))
                self._const_cache[key] = lo_hi
            raw = torch.cat([deviation_network.variance, beta_network.beta, beta_network.gamma])
            return torch.clamp(torch.exp(raw * 10.0), lo_hi[0], lo_hi[1])
        inv_s = deviation_network(torch.zeros([1, 3], device=self.device))[:, :1].clip(1e-6, 1e6)
        beta = beta_network.get_beta().clip(1e-6, 1e6)
        gamma = beta_network.get_gamma().clip(1e-6, 1e6)
        return torch.cat([inv_s.reshape(1), beta.reshape(1), gamma.reshape(1)])

    # ------------------------------------------------------------------ a13 render core
    def render_core(self, rays_o, rays_d, z_vals, sample_dist, udf_network, deviation_network,
                    beta_network=None, cos_anneal_ratio=None, background_rgb=None,
                    background_alpha=None, background_sampled_edge=None, flip_saturation=0.0):
        if background_alpha is not None:
            raise NotImplementedError("background_alpha branch is unreachable in the reference (SURVEY §0)")
        B, n = z_vals.shape
        if not torch.is_tensor(sample_dist):
            sample_dist = torch.tensor([sample_dist], dtype=torch.float32, device=self.device)
        dists, mid_z = ops.render_prep(z_vals, sample_dist)
        udf, grad = udf_network.udf_and_gradient(rays_o=rays_o, rays_d=rays_d, z=mid_z)

        # scalar networks (:466-472): inv_s = exp(10 variance), beta = exp(10 beta).clip(0, 1/beta_min),
        # gamma = exp(10 gamma), each clipped to [1e-6, 1e6]; autograd handles them
        scalars = self._scalars(deviation_network, beta_network)
        inv_s, beta, gamma = scalars[0:1].reshape(1, 1), scalars[1:2], scalars[2:3]

        cfg = dict(cos_anneal_ratio=-1.0 if cos_anneal_ratio is None else float(cos_anneal_ratio),
                   flip_saturation=float(flip_saturation), near_surface=float(self.near_surface),
                   sparse_scale=float(self.sparse_scale_factor), use_unbias=int(self.use_unbias_render),
                   use_norm_grad=int(self.use_norm_grad_for_cosine), alpha_type=self._alpha_type,
                   global_stats=bool(self.global_batch_stats))
        (weights, edge, depth, normals, gerr, gerr_ns, sparse, alpha, grad_flip, inside,
         grad_mag) = render_core_fn(udf, grad, scalars, rays_o, rays_d, mid_z, dists, B, n, cfg)
        if background_rgb is not None:
            edge = edge + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
        return {
            "udf": udf.view(B, n), "edge": edge, "weights": weights,
            "s_val": (1.0 / inv_s).expand(B * n, 1), "beta": 1.0 / beta, "gamma": gamma,
            "depth": depth, "gradient_error": gerr, "gradient_error_near_surface": gerr_ns,
            "normals": normals, "gradients": grad.view(B, n, 3), "gradients_flip": grad_flip,
            "inside_sphere": inside, "gradient_mag": grad_mag, "alpha": alpha, "mid_z_vals": mid_z,
            "dists": dists, "sparse_error": sparse,
        }

    # ------------------------------------------------------------------ a8 render
    def render(self, rays_o, rays_d, near, far, depth_scale, cos_anneal_ratio=None,
               perturb_overwrite=-1, background_rgb=None, flip_saturation=0, color_maps=None, pose=None,
               fx=None, fy=None, img_index=None, rays_uv=None):
        dev = self.device
        ops.poll_status(dev)                       # non-blocking NaN guard (see module docstring)
        rays_o = rays_o.to(dev, torch.float32).contiguous()
        rays_d = rays_d.to(dev, torch.float32).contiguous()
        B = len(rays_o)
        n0 = self.n_samples

        if not isinstance(near, torch.Tensor):
            # scalars: same fp32 arithmetic as the reference, evaluated on the host -> no sync
            key = ("nf", float(near), float(far), n0)
            near_t, far_t, sd_t = self._const_cache.get(key, (None, None, None))
            if near_t is None:
                n_c, f_c = torch.Tensor([near]).view(1, 1), torch.Tensor([far]).view(1, 1)
                sd_c = ((f_c - n_c) / n0).mean().reshape(1)
                near_t, far_t, sd_t = n_c.to(dev), f_c.to(dev), sd_c.to(dev)
                self._const_cache[key] = (near_t, far_t, sd_t)
            per_ray = False
        else:
            near_t = near.to(dev, torch.float32).reshape(-1, 1).contiguous()
            far_t = far.to(dev, torch.float32).reshape(-1, 1).contiguous()
            sd_t = ((far_t - near_t) / n0).mean().reshape(1)        # stays on the device (:704)
            per_ray = near_t.shape[0] > 1
            if per_ray and near_t.shape[0] != B:
                raise ValueError("near/far must be scalars, [1,1] or [B,1]")

        perturb = self.perturb if perturb_overwrite < 0 else perturb_overwrite
        t_rand = None
        if perturb > 0:
            if self.perturb_on_device:
                t_rand = torch.rand([B, 1], device=dev) - 0.5
            elif ops.graph_capturing():
                raise RuntimeError("render() under CUDA-graph capture: set renderer.perturb_on_device = True "
                                   "(a host-side draw would be frozen into the graph)")
            else:
                # one draw from the GLOBAL CPU generator, exactly like the reference (:719) -> same seeds
                t_rand = (torch.rand([B, 1]) - 0.5).to(dev, non_blocking=True)
        elif not per_ray:
            raise ValueError("perturb == 0 with scalar near/far: the reference crashes here too "
                             "(z stays [1,n]); pass near/far as [B,1] tensors")
        z_vals = ops.coarse_z(near_t, far_t, per_ray, self._linspace01(n0), t_rand, B, n0)

        n_samples = n0
        if self.n_importance > 0:
            with ops.nvtx_range("emap.importance_sample"):
                if self.upsampling_type == "classical":
                    z_vals = self.importance_sample(rays_o, rays_d, z_vals, sd_t)
                else:
                    z_vals = self.importance_sample_mix(rays_o, rays_d, z_vals, sd_t)
            n_samples = n0 + self.n_importance

        with ops.nvtx_range("emap.render_core"):
            r = self.render_core(rays_o, rays_d, z_vals, sd_t, self.udf_network, self.deviation_network,
                                 beta_network=self.beta_network, cos_anneal_ratio=cos_anneal_ratio,
                                 background_rgb=background_rgb, flip_saturation=flip_saturation)
        w = r["weights"]
        return {
            "udf": r["udf"], "edge": r["edge"],
            "weight_sum": w[:, :n_samples].sum(dim=-1, keepdim=True),
            "weight_sum_fg_bg": w.sum(dim=-1, keepdim=True),
            "depth": r["depth"] * depth_scale.to(dev), "variance": r["s_val"], "beta": r["beta"],
            "gamma": r["gamma"], "normals": r["normals"], "gradients": r["gradients"],
            "gradients_flip": r["gradients_flip"], "weights": w,
            "gradient_error": r["gradient_error"],
            "gradient_error_near_surface": r["gradient_error_near_surface"],
            "inside_sphere": r["inside_sphere"], "gradient_mag": r["gradient_mag"],
            "mid_z_vals": r["mid_z_vals"], "dists": r["dists"],
        }
