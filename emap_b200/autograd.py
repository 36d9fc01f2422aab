"""torch.autograd glue: exposes the CUDA kernels as differentiable functions of the nn.Parameters.

forward  : K0 fold (if parameters changed) + K1 / K1g
backward : K1b -- one fused backward for the pair (udf, d udf/dx): the cotangents (d_udf, d_grad)
           are pulled back to all 462,980 parameters (this includes the second-order terms that
           the reference obtains with autograd.grad(create_graph=True), udf_model.py:127-134).
Inputs x never receive gradients: in the reference's hot path the sample positions are detached
(importance sampling runs under no_grad, rays are data).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def _split_flat(flat: torch.Tensor, params):
    out, off = [], 0
    for p in params:
        n = p.numel()
        out.append(flat[off:off + n].view_as(p))
        off += n
    return out


class _UDFForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, rays_o, rays_d, z, want_pe, *params):
        net = module.packed()
        udf, pe = ops.udf_forward(net, module.prec_code, pts=x, rays_o=rays_o, rays_d=rays_d, z=z,
                                  want_pe=want_pe)
        ctx.module = module
        ctx.pts = (x, rays_o, rays_d, z)
        ctx.mark_non_differentiable(*([pe] if pe is not None else []))
        return (udf, pe) if want_pe else (udf,)

    @staticmethod
    def backward(ctx, d_udf, *unused):
        module = ctx.module
        x, ro, rd, z = ctx.pts
        net = module.packed()
        flat_grad = ops.udf_backward(net, module.prec_code, d_udf.contiguous(), None,
                                     pts=x, rays_o=ro, rays_d=rd, z=z, flat_params=net.flat)
        grads = _split_flat(flat_grad, module.flat_param_list())
        return (None, None, None, None, None, None, *grads)


class _UDFForwardGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, rays_o, rays_d, z, grad_mode_on, *params):
        net = module.packed()
        # shared-forward backward: when parameter gradients will be asked for, the reverse-mode forward also fills
        # the value rows of the backward's stashes.  needs_input_grad ignores torch.no_grad() (without the grad-mode
        # test every inference render() wrote -- and allocated -- the 4.4 GB stash too; found in the round-2 ncu
        # capture as 8.8 GB of stores by an "inference" launch), and grad mode is always OFF inside
        # Function.forward, so the caller samples it (udf_forward_grad_fn) and passes it in.
        stash = None
        if ops.shared_backward() and grad_mode_on and any(ctx.needs_input_grad[6:]):
            P = x.shape[0] if x is not None else z.numel()
            stash = ops.alloc_backward_stash(P, net.packed.device)
        udf, grad = ops.udf_forward_grad(net, module.prec_code, pts=x, rays_o=rays_o, rays_d=rays_d, z=z,
                                         stash=stash)
        ctx.set_materialize_grads(False)
        ctx.module = module
        ctx.pts = (x, rays_o, rays_d, z)
        ctx.stash = stash
        ctx.fold_id = getattr(net, "fold_id", None)
        return udf, grad

    @staticmethod
    def backward(ctx, d_udf, d_grad):
        if d_udf is None and d_grad is None:
            return (None,) * (6 + len(ctx.module.flat_param_list()))
        module = ctx.module
        x, ro, rd, z = ctx.pts
        net = module.packed()
        stash = ctx.stash
        if stash is not None and ctx.fold_id != getattr(net, "fold_id", None):
            stash = None            # parameters changed between forward and backward: recompute (dual forward)
        flat_grad = ops.udf_backward(net, module.prec_code,
                                     None if d_udf is None else d_udf.contiguous(),
                                     None if d_grad is None else d_grad.contiguous(),
                                     pts=x, rays_o=ro, rays_d=rd, z=z, flat_params=net.flat, stash=stash)
        ctx.stash = None
        grads = _split_flat(flat_grad, module.flat_param_list())
        return (None, None, None, None, None, None, *grads)


def _detach(t):
    return None if t is None else t.detach()


def udf_forward_fn(module, x=None, rays_o=None, rays_d=None, z=None, want_pe=False):
    params = module.flat_param_list()
    out = _UDFForward.apply(module, _detach(x), _detach(rays_o), _detach(rays_d), _detach(z), want_pe,
                            *params)
    return (out[0], out[1]) if want_pe else (out[0], None)


def udf_forward_grad_fn(module, x=None, rays_o=None, rays_d=None, z=None):
    params = module.flat_param_list()
    return _UDFForwardGrad.apply(module, _detach(x), _detach(rays_o), _detach(rays_d), _detach(z),
                                 torch.is_grad_enabled(), *params)
