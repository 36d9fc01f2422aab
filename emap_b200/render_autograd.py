"""Differentiable wrapper of the post-MLP render_core stage (K3 forward / backward kernels).

Differentiable outputs: weights[B,n], edge[B,1], depth[B,1], normals[B,3], gradient_error,
gradient_error_near_surface, sparse_error.  Differentiable inputs: udf[P], grad[P,3] (the MLP
outputs) and scalars = [inv_s, beta, gamma].  alpha / gradients_flip / inside_sphere / gradient_mag
are produced by the same kernel as plain (non-differentiable) tensors.
"""
from __future__ import annotations

import torch

from . import ops


class _RenderCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, udf, grad, scalars, rays_o, rays_d, mid_z, dists, B, n, cfg):
        (weights, alpha, grad_flip, inside, grad_mag, edge, depth, normals, reduced) = ops.render_core_fwd(
            rays_o, rays_d, mid_z, dists, udf, grad, scalars, B, n, cfg)
        if cfg.get("global_stats"):
            from .parallel import globalize_eikonal
            reduced = globalize_eikonal(reduced)       # exact full-batch eikonal means across ray shards
        ctx.set_materialize_grads(False)      # unused outputs arrive as None (a NULL cotangent), not as zero tensors
        ctx.save_for_backward(udf, grad, scalars, rays_o, rays_d, mid_z, dists, reduced)
        ctx.cfg, ctx.B, ctx.n = cfg, B, n
        gerr, gerr_ns, sparse = reduced[0], reduced[1], reduced[2]
        ctx.mark_non_differentiable(alpha, grad_flip, inside, grad_mag)
        return weights, edge, depth, normals, gerr, gerr_ns, sparse, alpha, grad_flip, inside, grad_mag

    @staticmethod
    def backward(ctx, d_w, d_edge, d_depth, d_normals, d_gerr, d_gerr_ns, d_sparse, *unused):
        udf, grad, scalars, rays_o, rays_d, mid_z, dists, reduced = ctx.saved_tensors
        d_udf, d_grad, d_scalars = ops.render_core_bwd(
            rays_o, rays_d, mid_z, dists, udf, grad, scalars, reduced, ctx.B, ctx.n, ctx.cfg,
            d_w, d_edge, d_depth, d_normals, d_gerr, d_gerr_ns, d_sparse)
        return d_udf, d_grad, d_scalars, None, None, None, None, None, None, None


def render_core_fn(udf, grad, scalars, rays_o, rays_d, mid_z, dists, B, n, cfg):
    return _RenderCore.apply(udf, grad, scalars, rays_o, rays_d, mid_z, dists, B, n, cfg)
