"""Thin functional layer over the C-ABI: owns device buffers (torch tensors), passes raw pointers.

Nothing here computes on the host; every function is one (or a few) kernel launches on the current
CUDA stream.  Used by the drop-in classes in udf_model.py / udf_renderer_blending.py.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _cabi as C


def net_dims(multires: int):
    pe = 3 + 6 * multires
    in_dim = [pe] + [256] * 8
    out_dim = [256, 256, 256, 256 - pe, 256, 256, 256, 256, 1]
    return in_dim, out_dim


class PackedNet:
    """Device-side image of one UDFNetwork: flat fp32 parameters + tensor-core operand pack."""

    def __init__(self, multires: int = 10, udf_type: str = "abs", scale: float = 1.0,
                 elem_type: str = "fp16", device="cuda"):
        self.desc = C.NetDesc(int(multires), C.UDF_TYPES[udf_type], float(scale),
                              0 if elem_type == "fp16" else 1)
        self.multires = int(multires)
        self.device = torch.device(device)
        L = C.lib()
        self.n_params = int(L.emap_flat_param_count(ctypes.byref(self.desc)))
        nbytes = int(L.emap_packed_size(ctypes.byref(self.desc)))
        if self.n_params == 0 or nbytes == 0:
            C.check(1)
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def fold(self, flat_params: torch.Tensor) -> None:
        """K0: W = g v/||v||, split/pack into tcgen05 operand images (once per optimizer step)."""
        flat_params = C.f32(flat_params)
        if flat_params.numel() != self.n_params:
            raise RuntimeError(f"flat parameter buffer has {flat_params.numel()} elements, "
                               f"expected {self.n_params}")
        C.check(C.lib().emap_wn_fold(ctypes.byref(self.desc), C.ptr(flat_params), C.ptr(self.packed),
                                     C.stream()))


def _points_args(pts, rays_o, rays_d, z):
    if pts is not None:
        pts = C.f32(pts)
        if pts.dim() != 2 or pts.shape[1] != 3:
            raise RuntimeError("pts must be [P,3]")
        return pts, None, None, None, 0, pts.shape[0]
    rays_o, rays_d, z = C.f32(rays_o), C.f32(rays_d), C.f32(z)
    B, n = z.shape
    if rays_o.shape != (B, 3) or rays_d.shape != (B, 3):
        raise RuntimeError("rays_o/rays_d must be [B,3] matching z [B,n]")
    return None, rays_o, rays_d, z, n, B * n


def udf_forward(net: PackedNet, precision: int, pts=None, rays_o=None, rays_d=None, z=None,
                want_pe: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    pts, ro, rd, zz, n, P = _points_args(pts, rays_o, rays_d, z)
    dev = net.packed.device
    udf = torch.empty(P, dtype=torch.float32, device=dev)
    pe = torch.empty(P, 3 + 6 * net.multires, dtype=torch.float32, device=dev) if want_pe else None
    C.check(C.lib().emap_udf_forward(ctypes.byref(net.desc), C.ptr(net.packed), precision, C.ptr(pts),
                                     C.ptr(ro), C.ptr(rd), C.ptr(zz), n, P, C.ptr(udf), C.ptr(pe),
                                     C.stream()))
    return udf, pe


def udf_forward_grad(net: PackedNet, precision: int, pts=None, rays_o=None, rays_d=None, z=None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    pts, ro, rd, zz, n, P = _points_args(pts, rays_o, rays_d, z)
    dev = net.packed.device
    udf = torch.empty(P, dtype=torch.float32, device=dev)
    grad = torch.empty(P, 3, dtype=torch.float32, device=dev)
    C.check(C.lib().emap_udf_forward_grad(ctypes.byref(net.desc), C.ptr(net.packed), precision,
                                          C.ptr(pts), C.ptr(ro), C.ptr(rd), C.ptr(zz), n, P,
                                          C.ptr(udf), C.ptr(grad), C.stream()))
    return udf, grad


def debug_mlp(net: PackedNet, precision: int, mode: int, pts: torch.Tensor):
    """Test hook: also returns the de-scaled accumulators of tile 0, [9,128,256]."""
    pts = C.f32(pts)
    P = pts.shape[0]
    dev = net.packed.device
    udf = torch.zeros(P, dtype=torch.float32, device=dev)
    grad = torch.zeros(P, 3, dtype=torch.float32, device=dev)
    dbg = torch.zeros(9, 128, 256, dtype=torch.float32, device=dev)
    C.check(C.lib().emap_debug_mlp(ctypes.byref(net.desc), C.ptr(net.packed), precision, mode,
                                   C.ptr(pts), P, C.ptr(udf), C.ptr(grad), C.ptr(dbg), C.stream()))
    return udf, grad, dbg


def positional_encoding(x: torch.Tensor, multires: int) -> torch.Tensor:
    """gamma(x) as computed by the MLP kernel's input stage (reference column order)."""
    x = C.f32(x.reshape(-1, 3))
    net = _pe_net(multires, x.device)
    _, pe = udf_forward(net, C.PREC_HALF, pts=x, want_pe=True)
    return pe


_PE_NETS = {}


def _pe_net(multires: int, device) -> PackedNet:
    key = (multires, str(device))
    if key not in _PE_NETS:
        net = PackedNet(multires, device=device)
        net.fold(torch.zeros(net.n_params, dtype=torch.float32, device=device) + 1.0)
        _PE_NETS[key] = net
    return _PE_NETS[key]


# ----------------------------------------------------------------------------- per-ray kernels
def coarse_z(near: torch.Tensor, far: torch.Tensor, per_ray: bool, lin: torch.Tensor,
             t_rand: Optional[torch.Tensor], B: int, n: int) -> torch.Tensor:
    z = torch.empty(B, n, dtype=torch.float32, device=lin.device)
    C.check(C.lib().emap_coarse_z(C.ptr(C.f32(near)), C.ptr(C.f32(far)), int(per_ray), C.ptr(lin),
                                  C.ptr(None if t_rand is None else C.f32(t_rand)), B, n, C.ptr(z),
                                  C.stream()))
    return z


def upsample_step(rays_o, rays_d, z_in, udf_in, z_add, udf_add, u, k, sample_dist, inv_s, beta, gamma,
                  mode=0, alpha_type=0, want_inds=False, want_weights=False):
    """Fused [merge pending samples] + [one up-sampling step].  Returns
    (z_cur[B,n+ka], udf_cur or None, z_new[B,k] or None, inds or None, weights or None)."""
    z_in = C.f32(z_in)
    B, n = z_in.shape
    dev = z_in.device
    ka = 0 if z_add is None else z_add.shape[1]
    z_out = udf_out = None
    if ka > 0:
        z_out = torch.empty(B, n + ka, dtype=torch.float32, device=dev)
        if udf_add is not None:
            udf_out = torch.empty(B, n + ka, dtype=torch.float32, device=dev)
    z_new = torch.empty(B, k, dtype=torch.float32, device=dev) if k > 0 else None
    inds = torch.empty(B, k, dtype=torch.int64, device=dev) if (k > 0 and want_inds) else None
    w = torch.empty(B, n + ka - 1, dtype=torch.float32, device=dev) if (k > 0 and want_weights) else None
    C.check(C.lib().emap_upsample_step(
        C.ptr(rays_o), C.ptr(rays_d), C.ptr(z_in), C.ptr(None if udf_in is None else C.f32(udf_in)), n,
        C.ptr(None if z_add is None else C.f32(z_add)), C.ptr(None if udf_add is None else C.f32(udf_add)),
        ka, C.ptr(z_out), C.ptr(udf_out), C.ptr(u), k, C.ptr(z_new), C.ptr(inds), C.ptr(w),
        C.ptr(sample_dist), B, float(inv_s), float(beta), float(gamma), int(mode), int(alpha_type),
        C.stream()))
    z_cur = z_out if ka > 0 else z_in
    udf_cur = udf_out if ka > 0 else udf_in
    return z_cur, udf_cur, z_new, inds, w


def sample_pdf_det(bins: torch.Tensor, weights: torch.Tensor, k: int):
    """sample_pdf(bins, weights, k, det=True) in isolation (udf_renderer_blending.py:69-109):
    the up-sampling kernel in mode 2 takes the weights as given.  Returns (samples, inds)."""
    bins, weights = C.f32(bins), C.f32(weights)
    B, n = bins.shape
    dev = bins.device
    u = torch.linspace(0.0 + 0.5 / k, 1.0 - 0.5 / k, steps=k).to(dev)
    sd = torch.zeros(1, dtype=torch.float32, device=dev)
    dummy = torch.zeros(B, 3, dtype=torch.float32, device=dev)
    _, _, z_new, inds, _ = upsample_step(dummy, dummy, bins, weights, None, None, u, k, sd, 1.0, 1.0, 1.0,
                                         mode=2, want_inds=True)
    return z_new, inds


def render_prep(z: torch.Tensor, sample_dist: torch.Tensor):
    z = C.f32(z)
    B, n = z.shape
    dists = torch.empty_like(z)
    mid = torch.empty_like(z)
    C.check(C.lib().emap_render_prep(C.ptr(z), C.ptr(sample_dist), B, n, C.ptr(dists), C.ptr(mid),
                                     C.stream()))
    return dists, mid


def render_core_fwd(rays_o, rays_d, mid_z, dists, udf, grad, scalars, B, n, cfg, want_alpha=True):
    dev = mid_z.device
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
    weights, grad_flip, inside, grad_mag = f(B, n), f(B, n, 3), f(B, n), f(B, n)
    alpha = f(B, n) if want_alpha else None
    edge, depth, normals = f(B, 1), f(B, 1), f(B, 3)
    partials = torch.empty(B, 5, dtype=torch.float64, device=dev)
    reduced = f(5)
    C.check(C.lib().emap_render_core_fwd(
        C.ptr(rays_o), C.ptr(rays_d), C.ptr(mid_z), C.ptr(dists), C.ptr(C.f32(udf)), C.ptr(C.f32(grad)),
        C.ptr(C.f32(scalars)), B, n, cfg["cos_anneal_ratio"], cfg["flip_saturation"],
        cfg["near_surface"], cfg["sparse_scale"], cfg["use_unbias"], cfg["use_norm_grad"],
        cfg["alpha_type"], C.ptr(weights), C.ptr(alpha), C.ptr(grad_flip), C.ptr(inside),
        C.ptr(grad_mag), C.ptr(edge), C.ptr(depth), C.ptr(normals), C.ptr(partials), C.ptr(reduced),
        C.stream()))
    return weights, alpha, grad_flip, inside, grad_mag, edge, depth, normals, reduced


def _opt(t):
    return None if t is None else C.f32(t)


def render_core_bwd(rays_o, rays_d, mid_z, dists, udf, grad, scalars, reduced, B, n, cfg, d_w, d_edge,
                    d_depth, d_normals, d_gerr, d_gerr_ns, d_sparse):
    dev = mid_z.device
    d_udf = torch.empty(B * n, dtype=torch.float32, device=dev)
    d_grad = torch.empty(B * n, 3, dtype=torch.float32, device=dev)
    partials = torch.empty(B, 3, dtype=torch.float64, device=dev)
    d_scalars = torch.empty(3, dtype=torch.float32, device=dev)
    C.check(C.lib().emap_render_core_bwd(
        C.ptr(rays_o), C.ptr(rays_d), C.ptr(mid_z), C.ptr(dists), C.ptr(C.f32(udf)), C.ptr(C.f32(grad)),
        C.ptr(C.f32(scalars)), C.ptr(reduced), B, n, cfg["cos_anneal_ratio"], cfg["flip_saturation"],
        cfg["near_surface"], cfg["sparse_scale"], cfg["use_unbias"], cfg["use_norm_grad"],
        cfg["alpha_type"], C.ptr(_opt(d_w)), C.ptr(_opt(d_edge)), C.ptr(_opt(d_depth)),
        C.ptr(_opt(d_normals)), C.ptr(_opt(d_gerr)), C.ptr(_opt(d_gerr_ns)), C.ptr(_opt(d_sparse)),
        C.ptr(d_udf), C.ptr(d_grad), C.ptr(partials), C.ptr(d_scalars), C.stream()))
    return d_udf, d_grad, d_scalars
