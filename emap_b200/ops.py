"""Thin functional layer over the C-ABI: owns device buffers (torch tensors), passes raw pointers.

Nothing here computes on the host; every function is one (or a few) kernel launches on the current
CUDA stream.  Used by the drop-in classes in udf_model.py / udf_renderer_blending.py.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _cabi as C


# ----------------------------------------------------------------------------- device status word
# The reference guards its hot path with NaN checks that drop into pdb (udf_renderer_blending.py:102-107,
# :346-351, :632-633) -- each a device->host sync.  Here the kernels OR bits into one int32 per device
# (include/emap_b200.h: EMAP_STATUS_*) and the host polls it lazily: poll_status() never blocks (it reads the
# copy enqueued by the previous poll once that has landed), check_status() synchronises.  Both raise
# FloatingPointError naming the stage and clear the word.
STATUS_BITS = {1: "NaN among the new z samples of an up-sampling step (sample_pdf / up_sample)",
               2: "NaN gradient_error in render_core",
               4: "non-finite parameter gradient out of the MLP backward"}
_STATUS = {}


class _Status:
    def __init__(self, dev):
        self.word = torch.zeros(1, dtype=torch.int32, device=dev)
        self.host = None                      # pinned landing buffer of the lazy poll, made on first use
        self.event = None


def _status(dev) -> _Status:
    dev = torch.device(dev)
    key = (dev.type, dev.index if (dev.index is not None or dev.type != "cuda") else torch.cuda.current_device())
    st = _STATUS.get(key)
    if st is None:
        st = _STATUS[key] = _Status(dev)
    return st


def status_word(dev) -> torch.Tensor:
    return _status(dev).word


def _raise_status(st: _Status, bits: int):
    st.word.zero_()
    st.event = None
    what = "; ".join(msg for b, msg in STATUS_BITS.items() if bits & b) or f"status {bits}"
    raise FloatingPointError("emap_b200: " + what)


def poll_status(dev) -> None:
    """Non-blocking: raise if the copy enqueued by the previous poll has landed with a bit set, then enqueue
    the next copy (4 bytes, pinned, on the current stream)."""
    st = _status(dev)
    if torch.cuda.is_current_stream_capturing():
        return
    if st.event is not None and st.event.query():
        bits = int(st.host[0])
        st.event = None
        if bits:
            _raise_status(st, bits)
    if st.event is None:
        if st.host is None:
            st.host = torch.zeros(1, dtype=torch.int32).pin_memory()
        st.host.copy_(st.word, non_blocking=True)
        st.event = torch.cuda.Event()
        st.event.record()


def check_status(dev) -> None:
    """Blocking check (one 4-byte device->host read)."""
    st = _status(dev)
    bits = int(st.word.item())
    if bits:
        _raise_status(st, bits)


# set by graph.GraphedStep while it warms up / captures an iteration: host-side draws and other things that a CUDA
# graph would freeze are refused already in the warm-up, before any capture has begun
graph_prepare = False


def graph_capturing() -> bool:
    return graph_prepare or torch.cuda.is_current_stream_capturing()


# ----------------------------------------------------------------------------- NVTX ranges
# EMAP_NVTX=1 brackets the stages of the path (render, importance sampling, render_core, the backward's stages)
# with NVTX ranges for Nsight timelines; off by default (the reference has no tracing at all).
_NVTX = os.environ.get("EMAP_NVTX") == "1"


class nvtx_range:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


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

    def backward_net(self) -> "PackedNet":
        """The operand pack the backward kernels use: fp16 images.  A bf16 network keeps a second, fp16 pack of
        the same parameters for it (re-folded when the parameters change): bf16 forward, fp16 backward."""
        if self.desc.elem_type == 0:
            return self
        bn = getattr(self, "_bwd_net", None)
        if bn is None:
            bn = PackedNet(self.multires, [k for k, v in C.UDF_TYPES.items() if v == self.desc.udf_type][0],
                           float(self.desc.scale), "fp16", self.device)
            self._bwd_net = bn
        if getattr(bn, "src_fold_id", None) != self.fold_id:
            bn.fold(self.flat)
            bn.src_fold_id = self.fold_id
        return bn

    def fold(self, flat_params: torch.Tensor) -> None:
        """K0: W = g v/||v||, split/pack into tcgen05 operand images (once per optimizer step)."""
        flat_params = C.f32(flat_params)
        if flat_params.numel() != self.n_params:
            raise RuntimeError(f"flat parameter buffer has {flat_params.numel()} elements, "
                               f"expected {self.n_params}")
        C.check(C.lib().emap_wn_fold(ctypes.byref(self.desc), C.ptr(flat_params), C.ptr(self.packed),
                                     C.stream()))
        self.flat = flat_params                         # kept for the backward (bias / g / v reads)
        self.fold_id = getattr(self, "fold_id", 0) + 1


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


# How (udf, d udf/dx) is evaluated: "reverse" = K1r, value forward + adjoint sweep in one kernel (mlp_rg.cu;
# what autograd does, half the tensor-core work of forward mode -- the default since it was validated on
# B200); "forward" = K1g, forward-mode tangent rows (mlp_tc.cu MODE 1; kept as an independent cross-check).
DEFAULT_GRAD_MODE = os.environ.get("EMAP_GRAD_MODE", "reverse")
_GRAD_MODE = DEFAULT_GRAD_MODE
_RG_SCRATCH = {}


# How the backward obtains the dual activations U_l = (h_l ; hdot_l): "shared" (default) = the training
# forward (K1r) writes the value rows while it has them in registers and the backward only adds the tangent
# rows (emap_bwd_tangent_forward, about half the work); "dual" = re-run the dual forward (cross-check, and
# what grad mode "forward" uses).  "shared" needs grad mode "reverse".
DEFAULT_BWD_MODE = os.environ.get("EMAP_BWD_STASH", "shared")
_BWD_MODE = DEFAULT_BWD_MODE


def set_backward_mode(mode: str) -> None:
    global _BWD_MODE
    if mode not in ("dual", "shared"):
        raise ValueError("backward mode must be 'dual' or 'shared'")
    _BWD_MODE = mode


def shared_backward() -> bool:
    return _BWD_MODE == "shared" and _GRAD_MODE == "reverse"


def alloc_backward_stash(P: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """(st_u0 [2P,64], st_u [8,2P,256]) fp16 -- the layout emap_bwd_dual_forward documents."""
    return (torch.empty(2 * P, 64, dtype=torch.float16, device=device),
            torch.empty(8, 2 * P, 256, dtype=torch.float16, device=device))


def set_grad_mode(mode: str) -> None:
    global _GRAD_MODE
    if mode not in ("forward", "reverse"):
        raise ValueError("grad mode must be 'forward' or 'reverse'")
    _GRAD_MODE = mode


def get_grad_mode() -> str:
    return _GRAD_MODE


def _rg_scratch(dev: torch.device) -> torch.Tensor:
    """K1r's sigma scratch (448 KiB per SM), one per device; kernels on one stream reuse it in order."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    buf = _RG_SCRATCH.get(key)
    if buf is None:
        with torch.cuda.device(dev):
            nbytes = int(C.lib().emap_rgrad_scratch_bytes())
        buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _RG_SCRATCH[key] = buf
    return buf


_BWD_WS = {}


def _bwd_workspace(dev: torch.device) -> torch.Tensor:
    """Partial-gradient workspace of the backward's weight-gradient stage (~290 MB), one per device; calls on
    one stream reuse it in order."""
    dev = torch.device(dev)
    key = (dev.type, dev.index if (dev.index is not None or dev.type != "cuda") else torch.cuda.current_device())
    buf = _BWD_WS.get(key)
    if buf is None:
        if dev.type == "cuda":
            with torch.cuda.device(dev):
                nbytes = int(C.lib().emap_bwd_workspace_bytes())
        else:
            nbytes = int(C.lib().emap_bwd_workspace_bytes())
        buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        _BWD_WS[key] = buf
    return buf


def udf_forward_grad(net: PackedNet, precision: int, pts=None, rays_o=None, rays_d=None, z=None,
                     mode: Optional[str] = None, stash=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """stash = (st_u0, st_u) from alloc_backward_stash (reverse mode only): the forward also writes the
    value rows of the backward's stashes."""
    pts, ro, rd, zz, n, P = _points_args(pts, rays_o, rays_d, z)
    dev = net.packed.device
    udf = torch.empty(P, dtype=torch.float32, device=dev)
    grad = torch.empty(P, 3, dtype=torch.float32, device=dev)
    if (mode or _GRAD_MODE) == "reverse":
        scratch = _rg_scratch(dev)
        su0, su = (None, None) if stash is None else stash
        if stash is not None and (su0.shape != (2 * P, 64) or su.shape != (8, 2 * P, 256)
                                  or su0.dtype != torch.float16 or su.dtype != torch.float16):
            raise RuntimeError("backward stash must be (fp16 [2P,64], fp16 [8,2P,256])")
        C.check(C.lib().emap_udf_forward_grad_rev(ctypes.byref(net.desc), C.ptr(net.packed), precision,
                                                  C.ptr(pts), C.ptr(ro), C.ptr(rd), C.ptr(zz), n, P,
                                                  C.ptr(udf), C.ptr(grad), C.ptr(scratch), scratch.numel(),
                                                  C.ptr(su0), C.ptr(su), C.stream()))
        return udf, grad
    if stash is not None:
        raise RuntimeError("a backward stash can only be filled by the reverse-mode forward (K1r)")
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


def debug_rgrad(net: PackedNet, precision: int, pts: torch.Tensor):
    """Test hook: K1r plus the accumulators of tile 0 after each MMA step, [16,128,256]."""
    pts = C.f32(pts)
    P = pts.shape[0]
    dev = net.packed.device
    udf = torch.zeros(P, dtype=torch.float32, device=dev)
    grad = torch.zeros(P, 3, dtype=torch.float32, device=dev)
    dbg = torch.zeros(16, 128, 256, dtype=torch.float32, device=dev)
    scratch = _rg_scratch(dev)
    C.check(C.lib().emap_debug_rgrad(ctypes.byref(net.desc), C.ptr(net.packed), precision, C.ptr(pts), None,
                                     None, None, 0, P, C.ptr(udf), C.ptr(grad), C.ptr(scratch),
                                     scratch.numel(), C.ptr(dbg), C.stream()))
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
                  mode=0, alpha_type=0, want_inds=False, want_weights=False, gamma_dev=None):
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
        C.ptr(sample_dist), B, float(inv_s), float(beta), float(gamma),
        C.ptr(None if gamma_dev is None else C.f32(gamma_dev)), int(mode), int(alpha_type),
        C.ptr(status_word(dev)), C.stream()))
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
        C.ptr(status_word(dev)), C.stream()))
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


# ----------------------------------------------------------------------------- K1b backward
def _weff_views(net: PackedNet):
    """fp32 views of W_eff (layer 0..8) inside the packed buffer + cached fp16 GEMM operands."""
    if getattr(net, "_offsets", None) is None:
        arr = (ctypes.c_uint32 * 10)()
        C.check(C.lib().emap_packed_offsets(ctypes.byref(net.desc), arr))
        net._offsets = list(arr)
    in_dim, out_dim = net_dims(net.multires)
    W = []
    for l in range(9):
        off = net._offsets[1 + l]
        n = in_dim[l] * out_dim[l]
        W.append(net.packed[off:off + 4 * n].view(torch.float32).view(out_dim[l], in_dim[l]))
    return W, in_dim, out_dim


def udf_backward(net: PackedNet, precision: int, d_udf: Optional[torch.Tensor],
                 d_grad: Optional[torch.Tensor], pts=None, rays_o=None, rays_d=None, z=None,
                 flat_params: Optional[torch.Tensor] = None, stash=None) -> torch.Tensor:
    """Pull the cotangents (d_udf[P], d_grad[P,3]) back to the flat parameter gradient.

    fp16 operand images and stashes, fp32 accumulation, loss-scaled on the device
    (emap_bwd_cotangent_scales): [tangent forward | dual forward] -> output-layer pull-back (with dW_8, db_8)
    -> reverse sweep -> the weight-gradient contractions dW_l = A_l^T U_l with the bias sums (all tcgen05 kernels)
    -> fixed-order sum of the partials + weight-norm backward, which removes the loss scale and flags non-finite
    gradients.  No library GEMM anywhere."""
    L = C.lib()
    net = net.backward_net()
    pts, ro, rd, zz, n, P = _points_args(pts, rays_o, rays_d, z)
    dev = net.packed.device
    if flat_params is None:
        raise RuntimeError("udf_backward needs the flat parameter buffer")
    flat_params = C.f32(flat_params)
    W, in_dim, out_dim = _weff_views(net)
    st = C.stream()
    desc = ctypes.byref(net.desc)
    h16 = lambda *s: torch.empty(*s, dtype=torch.float16, device=dev)  # noqa: E731
    boff, off = [], 0
    for l in range(9):
        boff.append(off)
        off += out_dim[l] * (2 + in_dim[l])
    d_udf = None if d_udf is None else C.f32(d_udf)
    d_grad = None if d_grad is None else C.f32(d_grad)

    scales = None
    if os.environ.get("EMAP_BWD_NOSCALE") != "1":          # (diagnostic switch: what round 1 did, raw cotangents)
        scales = torch.empty(8, dtype=torch.float32, device=dev)
        C.check(L.emap_bwd_cotangent_scales(C.ptr(d_udf), C.ptr(d_grad), P, C.ptr(scales), st))
    st_a = h16(8, 2 * P, 256)
    if stash is not None:
        # value rows written by the training forward (K1r): add the tangent rows only
        st_u0, st_u = stash
        with nvtx_range("emap.bwd.tangent_forward"):
            C.check(L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(pts), C.ptr(ro), C.ptr(rd), C.ptr(zz),
                                               n, P, C.ptr(d_grad), C.ptr(scales), C.ptr(st_u0), C.ptr(st_u), st))
    else:
        st_u0, st_u = h16(2 * P, 64), h16(8, 2 * P, 256)
        C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), C.PREC_HALF, C.ptr(pts), C.ptr(ro), C.ptr(rd),
                                        C.ptr(zz), n, P, C.ptr(d_grad), C.ptr(scales), C.ptr(st_u0),
                                        C.ptr(st_u), st))
    coef = torch.empty(2 * P, dtype=torch.float32, device=dev)
    ws = _bwd_workspace(dev)
    C.check(L.emap_bwd_top(desc, C.ptr(st_u[7]), C.ptr(W[8].reshape(-1)), C.ptr(flat_params[boff[8]:boff[8] + 1]),
                           C.ptr(d_udf), C.ptr(scales), P, C.ptr(coef), C.ptr(ws), st))
    with nvtx_range("emap.bwd.reverse_sweep"):
        C.check(L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(st_u), C.ptr(st_a), P, st))
    # dW_l = A_l^T U_l and db_l on the tensor cores (mlp_dw.cu), per-CTA partials; summed in a fixed order by the
    # final stage, which also undoes the PE column order, applies the weight-norm backward and removes the scale
    with nvtx_range("emap.bwd.weight_grads"):
        n_parts = int(L.emap_bwd_weight_grads(desc, C.ptr(st_a), C.ptr(st_u0), C.ptr(st_u), P, C.ptr(ws),
                                              ws.numel(), st))
    if n_parts <= 0:
        C.check(1)
    # (+ spare tail: parallel.FlatGradAllReduce parks the few foreign gradients there and all-reduces in place)
    flat_grad = torch.empty(flat_params.numel() + 32, dtype=torch.float32, device=dev)[:flat_params.numel()]
    C.check(L.emap_bwd_finish(desc, C.ptr(flat_params), C.ptr(ws), n_parts, C.ptr(scales), C.ptr(flat_grad),
                              C.ptr(status_word(dev)), st))
    return flat_grad


# ----------------------------------------------------------------------------- SURVEY §8f rows 1-2
def null_direction(grad_ld: torch.Tensor) -> torch.Tensor:
    """[M,S,3] stacked gradients -> [M,3] unit right-singular vector of the smallest singular value
    (extract_pointcloud.py:86-89: svd + vh[:, -1, :] + F.normalize), one thread per voxel."""
    grad_ld = C.f32(grad_ld)
    if grad_ld.dim() != 3 or grad_ld.shape[2] != 3:
        raise RuntimeError("grad_ld must be [M,S,3]")
    M, S, _ = grad_ld.shape
    out = torch.empty(M, 3, dtype=torch.float32, device=grad_ld.device)
    if M == 0:
        return out
    C.check(C.lib().emap_null_direction(C.ptr(grad_ld), M, S, C.ptr(out), C.stream()))
    return out


def rays_from_pixels(pixels_x, pixels_y, edge_img, intr_inv, pose):
    """pixels [B] int64 (device) of one image -> dict of the per-ray tensors of
    gen_random_rays_patches_at (dataset.py:268-305).  intr_inv / pose: [4,4] tensors (any device)."""
    if pixels_x.dtype != torch.int64 or pixels_y.dtype != torch.int64:
        raise RuntimeError("pixels must be int64")
    dev = pixels_x.device
    edge_img = C.f32(edge_img)
    H, W = int(edge_img.shape[0]), int(edge_img.shape[1])
    B = pixels_x.numel()
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
    rays_o, rays_v, edge, ndc, p_cam, ds = f(B, 3), f(B, 3), f(B, 1), f(B, 2), f(B, 3), f(B, 1)
    if B == 0:
        return {"rays_o": rays_o, "rays_v": rays_v, "edge": edge, "rays_ndc_uv": ndc,
                "rays_norm_XYZ_cam": p_cam, "depth_scale": ds}
    kinv = (ctypes.c_float * 9)(*[float(v) for v in intr_inv[:3, :3].reshape(-1).tolist()])
    pm = (ctypes.c_float * 16)(*[float(v) for v in pose.reshape(-1).tolist()])
    C.check(C.lib().emap_rays_from_pixels(C.ptr(pixels_x.contiguous()), C.ptr(pixels_y.contiguous()),
                                          C.ptr(edge_img), H, W, kinv, pm, B, C.ptr(rays_o), C.ptr(rays_v),
                                          C.ptr(edge), C.ptr(ndc), C.ptr(p_cam), C.ptr(ds), C.stream()))
    return {"rays_o": rays_o, "rays_v": rays_v, "edge": edge, "rays_ndc_uv": ndc,
            "rays_norm_XYZ_cam": p_cam, "depth_scale": ds}
