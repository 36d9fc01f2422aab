"""CPU oracle for the EMAP volume-rendering hot path.  TEST INFRASTRUCTURE ONLY.

This is a from-scratch *restatement* (plain torch on the CPU, fp32 by default,
fp64 on request) of the algorithm the reference implements in

    /root/reference/src/models/embedder.py              (positional encoding)
    /root/reference/src/models/udf_model.py             (UDFNetwork & scalar nets)
    /root/reference/src/models/udf_renderer_blending.py (sampling + compositing)

It exists so that the CUDA product in ``emap_b200/`` has something to be checked
against on a box where ``/root/reference`` does not exist (the GPU box).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product never does: the product
path fails loudly when the CUDA library is missing.

Pinning: the reference ships no tests, golden vectors or fixtures (SURVEY §4,
§8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced
in the build container by importing the reference modules read-only
(``tests/golden/make_golden.py``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` asserts this file reproduces every one of those
fixtures (bit-exact where the op order is identical, otherwise <= 2e-6).

Everything is functional: networks are a ``UDFParams`` bundle of plain tensors,
so autograd through the oracle yields the parameter gradients the CUDA backward
is compared with.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SOFTPLUS_BETA = 100.0  # udf_model.py:78  nn.Softplus(beta=100) (threshold 20)


# --------------------------------------------------------------------------- #
# parameters
# --------------------------------------------------------------------------- #
@dataclass
class UDFParams:
    """Weight-normed MLP parameters, one entry per nn.Linear (udf_model.py:39-76).

    v[l]: [out,in] direction ("original1"), g[l]: [out,1] magnitude ("original0"),
    b[l]: [out].  Effective weight  W_l = g_l * v_l / ||v_l||_row  (torch
    weight_norm, dim=0).
    """

    v: List[torch.Tensor]
    g: List[torch.Tensor]
    b: List[torch.Tensor]
    multires: int = 10
    skip_in: Tuple[int, ...] = (4,)
    scale: float = 1.0
    udf_type: str = "abs"

    @property
    def n_linear(self) -> int:
        return len(self.v)

    def tensors(self) -> List[torch.Tensor]:
        out: List[torch.Tensor] = []
        for l in range(self.n_linear):
            out += [self.b[l], self.g[l], self.v[l]]  # nn.Module.parameters() order
        return out

    def to(self, dtype) -> "UDFParams":
        return UDFParams(
            [t.detach().to(dtype) for t in self.v],
            [t.detach().to(dtype) for t in self.g],
            [t.detach().to(dtype) for t in self.b],
            self.multires, self.skip_in, self.scale, self.udf_type,
        )

    def requires_grad_(self, flag: bool = True) -> "UDFParams":
        for t in self.tensors():
            t.requires_grad_(flag)
        return self

    @staticmethod
    def from_state_dict(sd: Dict[str, torch.Tensor], multires=10, skip_in=(4,), scale=1.0,
                        udf_type="abs") -> "UDFParams":
        """Accepts the reference checkpoint key names (``linN.bias``,
        ``linN.parametrizations.weight.original0/1``)."""
        n = 0
        while f"lin{n}.bias" in sd:
            n += 1
        v = [sd[f"lin{l}.parametrizations.weight.original1"].detach().clone() for l in range(n)]
        g = [sd[f"lin{l}.parametrizations.weight.original0"].detach().clone() for l in range(n)]
        b = [sd[f"lin{l}.bias"].detach().clone() for l in range(n)]
        return UDFParams(v, g, b, multires, tuple(skip_in), scale, udf_type)


@dataclass
class ScalarParams:
    """SingleVarianceNetwork / BetaNetwork scalars (udf_model.py:212-286)."""

    variance: torch.Tensor  # [1]
    beta: torch.Tensor      # [1]
    gamma: torch.Tensor     # [1]
    beta_min: float = 5e-5

    def inv_s(self) -> torch.Tensor:
        # udf_model.py:226 exp(10*variance); clip at udf_renderer_blending.py:466
        return torch.exp(self.variance * 10.0).clip(1e-6, 1e6)

    def beta_val(self) -> torch.Tensor:
        # udf_model.py:259 then udf_renderer_blending.py:471
        return torch.exp(self.beta * 10).clip(0, 1.0 / self.beta_min).clip(1e-6, 1e6)

    def gamma_val(self) -> torch.Tensor:
        # udf_model.py:262 then udf_renderer_blending.py:472
        return torch.exp(self.gamma * 10).clip(1e-6, 1e6)


def geometric_init(d_in=3, d_out=1, d_hidden=256, n_layers=8, skip_in=(4,), multires=10,
                   bias=0.5, generator: Optional[torch.Generator] = None) -> UDFParams:
    """Sphere initialisation of udf_model.py:39-76, stated on plain tensors.

    (Not RNG-stream compatible with the reference constructor -- fixtures carry
    the reference's own state_dict -- but the same distribution.)
    """
    pe = d_in * (1 + 2 * multires) if multires > 0 else d_in
    dims = [pe] + [d_hidden] * n_layers + [d_out]
    v, g, b = [], [], []
    for l in range(len(dims) - 1):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        w = torch.empty(out_dim, dims[l])
        bb = torch.zeros(out_dim)
        std = math.sqrt(2) / math.sqrt(out_dim)
        if l == len(dims) - 2:
            w.normal_(math.sqrt(math.pi) / math.sqrt(dims[l]), 1e-4, generator=generator)
            bb.fill_(-bias)
        elif multires > 0 and l == 0:
            w.zero_()
            w[:, :3].normal_(0.0, std, generator=generator)
        elif multires > 0 and l in skip_in:
            w.normal_(0.0, std, generator=generator)
            w[:, -(dims[0] - 3):] = 0.0
        else:
            w.normal_(0.0, std, generator=generator)
        # weight_norm parametrisation: g = ||w||_row, v = w
        g.append(w.norm(dim=1, keepdim=True))
        v.append(w.clone())
        b.append(bb)
    return UDFParams(v, g, b, multires, tuple(skip_in))


# --------------------------------------------------------------------------- #
# a1  positional encoding          embedder.py:5-53
# --------------------------------------------------------------------------- #
def posenc(x: torch.Tensor, multires: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (embedder.py:21-35)."""
    if multires <= 0:
        return x
    feats = [x]
    for j in range(multires):
        f = float(2.0 ** j)  # 2**linspace(0, L-1, L) is exactly 1,2,4,...  [SURVEY a1]
        feats.append(torch.sin(x * f))
        feats.append(torch.cos(x * f))
    return torch.cat(feats, dim=-1)


# --------------------------------------------------------------------------- #
# a3-a5  UDF MLP                   udf_model.py:90-135
# --------------------------------------------------------------------------- #
def effective_weights(p: UDFParams) -> List[torch.Tensor]:
    """W_l = g * v / ||v||  per output row (torch weight_norm dim=0; udf_model.py:74)."""
    return [torch._weight_norm(p.v[l], p.g[l], 0) for l in range(p.n_linear)]


def udf_forward(p: UDFParams, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """UDFNetwork.forward (udf_model.py:90-110): returns (out[P,d_out], PE[P,63])."""
    x = x * p.scale
    e = posenc(x, p.multires)
    W = effective_weights(p)
    h = e
    last = p.n_linear - 1
    for l in range(p.n_linear):
        if l in p.skip_in:
            h = torch.cat([h, e], dim=1) / math.sqrt(2)
        h = F.linear(h, W[l], p.b[l])
        if l < last:
            h = F.softplus(h, beta=SOFTPLUS_BETA)
    first = h[:, :1]
    if p.udf_type == "abs":
        first = first.abs()
    elif p.udf_type == "square":
        first = first ** 2
    return torch.cat([first / p.scale, h[:, 1:]], dim=-1), e


def udf_gradient(p: UDFParams, x: torch.Tensor, create_graph: bool = True) -> torch.Tensor:
    """UDFNetwork.gradient (udf_model.py:121-135): d udf / d x, [P,3]."""
    x = x.detach().requires_grad_(True)
    with torch.enable_grad():
        y = udf_forward(p, x)[0][:, :1]
        (gx,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=create_graph,
                                    retain_graph=True)
    return gx


# --------------------------------------------------------------------------- #
# a11  density helpers            udf_renderer_blending.py:155-170, 379-416
# --------------------------------------------------------------------------- #
def udf2logistic(udf, inv_s, gamma=20.0, abs_cos_val=1.0):
    """udf_renderer_blending.py:163-170 (cos_anneal branch is never taken by callers)."""
    ex = torch.exp(-inv_s * udf)
    return abs_cos_val * inv_s * ex / (1 + ex) ** 2 * gamma


def sdf2alpha_numerical(sdf, true_cos, dists, inv_s, cos_anneal_ratio=None):
    """udf_renderer_blending.py:384-411 ("numerical" type)."""
    if cos_anneal_ratio is not None:
        iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio)
                     + F.relu(-true_cos) * cos_anneal_ratio)
    else:
        iter_cos = true_cos
    nxt = sdf + iter_cos * dists * 0.5
    prv = sdf - iter_cos * dists * 0.5
    prev_cdf = torch.sigmoid(prv * inv_s)
    next_cdf = torch.sigmoid(nxt * inv_s)
    return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)


def sdf2alpha_theorical(sdf, true_cos, dists, inv_s, cos_anneal_ratio=None):
    """udf_renderer_blending.py:412-414."""
    if cos_anneal_ratio is not None:
        iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio)
                     + F.relu(-true_cos) * cos_anneal_ratio)
    else:
        iter_cos = true_cos
    raw = iter_cos.abs() * inv_s * (1 - torch.sigmoid(sdf * inv_s))
    return 1.0 - torch.exp(-F.relu(raw) * dists)


def _excl_cumprod(x: torch.Tensor) -> torch.Tensor:
    """cumprod([1, x_0, x_1, ...])[:-1]  -- exclusive transmittance scan."""
    ones = torch.ones_like(x[:, :1])
    return torch.cumprod(torch.cat([ones, x], dim=-1), dim=-1)[:, :-1]


# --------------------------------------------------------------------------- #
# a11b  inverse-CDF resampling     udf_renderer_blending.py:69-109
# --------------------------------------------------------------------------- #
def sample_pdf_det(bins: torch.Tensor, weights: torch.Tensor, k: int,
                   return_inds: bool = False):
    """Deterministic branch of sample_pdf (det=True): bins [B,n], weights [B,n-1] -> [B,k]."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)  # [B,n]
    u = torch.linspace(0.0 + 0.5 / k, 1.0 - 0.5 / k, steps=k, dtype=cdf.dtype).to(cdf.device)
    u = u.expand(cdf.shape[0], k).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp_min(0)
    above = inds.clamp_max(cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = cdf.gather(1, below), cdf.gather(1, above)
    bin_lo, bin_hi = bins.gather(1, below), bins.gather(1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_lo) / denom
    samples = bin_lo + t * (bin_hi - bin_lo)
    if return_inds:
        return samples, inds
    return samples


# --------------------------------------------------------------------------- #
# a10  one up-sampling step        udf_renderer_blending.py:228-353
# --------------------------------------------------------------------------- #
def up_sample_unbias(rays_o, rays_d, z, udf, sample_dist, k, inv_s, beta, gamma,
                     return_aux: bool = False, sdf2alpha_type: str = "numerical"):
    B, n = z.shape
    alpha2 = sdf2alpha_numerical if sdf2alpha_type == "numerical" else sdf2alpha_theorical
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)               # :250

    udf = udf.reshape(B, n)
    last = torch.full_like(z[:, :1], float(sample_dist))
    dists_raw = torch.cat([z[:, 1:] - z[:, :-1], last], -1)               # :254-263

    prev_u, next_u = udf[:, :-1], udf[:, 1:]
    prev_z, next_z = z[:, :-1], z[:, 1:]
    mid_udf = (prev_u + next_u) * 0.5
    dists = next_z - prev_z

    true_cos = (next_u - prev_u) / (next_z - prev_z + 1e-5)               # :279
    cos_val = -1 * torch.abs(true_cos)
    prev_cos = torch.cat([torch.zeros_like(cos_val[:, :1]), cos_val[:, :-1]], -1)
    cos_val = torch.minimum(prev_cos, cos_val)                            # :284-288
    cos_val = cos_val.clip(-1e3, 0.0) * inside                            # :290

    vis_mask = (true_cos < 0.05).to(z.dtype)                              # :293-296
    vis_mask = torch.cat([torch.ones_like(vis_mask[:, :1]), vis_mask], -1)

    raw_occ = udf2logistic(udf, beta, 1.0, 1.0)                           # :303
    alpha_occ = 1.0 - torch.exp(-F.relu(raw_occ) * gamma * dists_raw)     # :305
    vis_prob = _excl_cumprod((1.0 - alpha_occ + vis_mask).clip(0, 1) + 1e-7)  # :308-319

    signs_prob = vis_prob[:, :-1]
    a_plus = alpha2(mid_udf, cos_val, dists, inv_s)                       # :327-330
    a_minus = alpha2(-mid_udf, cos_val, dists, inv_s)
    alpha = a_plus * signs_prob + a_minus * (1 - signs_prob)
    weights = alpha * _excl_cumprod(1.0 - alpha + 1e-7)                   # :334-343
    z_new, inds = sample_pdf_det(z, weights, k, return_inds=True)         # :344
    if return_aux:
        return z_new, inds, weights
    return z_new


def up_sample_no_occ_aware(rays_o, rays_d, z, udf, sample_dist, k, inv_s, beta, gamma,
                           return_aux: bool = False, sdf2alpha_type: str = "numerical"):
    """udf_renderer_blending.py:920-975."""
    B, n = z.shape
    udf = udf.reshape(B, n)
    last = torch.full_like(z[:, :1], float(sample_dist))
    dists = torch.cat([z[:, 1:] - z[:, :-1], last], -1)
    raw_occ = udf2logistic(udf, beta, 1.0, 1.0)
    alpha_occ = 1.0 - torch.exp(-F.relu(raw_occ) * gamma * dists)
    z_new, inds = sample_pdf_det(z, alpha_occ[:, :-1], k, return_inds=True)
    if return_aux:
        return z_new, inds, alpha_occ[:, :-1]
    return z_new


# --------------------------------------------------------------------------- #
# a12  sorted merge                udf_renderer_blending.py:355-377
# --------------------------------------------------------------------------- #
def cat_z_vals(p: UDFParams, rays_o, rays_d, z, z_new, udf, last: bool):
    B, n = z.shape
    k = z_new.shape[1]
    zz, index = torch.sort(torch.cat([z, z_new], dim=-1), dim=-1)
    if not last:
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_new[..., :, None]
        new_udf = udf_forward(p, pts.reshape(-1, 3))[0][:, 0].reshape(B, k)
        udf = torch.cat([udf, new_udf], dim=-1).gather(1, index)
    return zz, udf


# --------------------------------------------------------------------------- #
# a9  hierarchical sampling        udf_renderer_blending.py:802-841
# --------------------------------------------------------------------------- #
def upsample_schedule(step: int, up_sample_steps: int) -> Tuple[float, float, float]:
    """(inv_s, beta, gamma) for step i (udf_renderer_blending.py:826-830)."""
    inv_s = 64.0 * 2 ** step
    beta = 64.0 * 2 ** (step + 1)
    gamma = float(min(max(20 * 2 ** (up_sample_steps - step), 20), 320))
    return inv_s, beta, gamma


@torch.no_grad()
def importance_sample(p: UDFParams, rays_o, rays_d, z, sample_dist, n_importance,
                      up_sample_steps, use_unbias_render: bool = True, trace: Optional[list] = None,
                      sdf2alpha_type: str = "numerical"):
    B, n0 = z.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]
    udf = udf_forward(p, pts.reshape(-1, 3))[0][:, 0].reshape(B, n0)
    k = n_importance // up_sample_steps
    step_fn = up_sample_unbias if use_unbias_render else up_sample_no_occ_aware
    for i in range(up_sample_steps):
        inv_s, beta, gamma = upsample_schedule(i, up_sample_steps)
        z_new, inds, w = step_fn(rays_o, rays_d, z, udf, sample_dist, k, inv_s, beta, gamma,
                                 return_aux=True, sdf2alpha_type=sdf2alpha_type)
        if trace is not None:
            trace.append({"z": z.clone(), "udf": udf.clone(), "z_new": z_new.clone(),
                          "inds": inds.clone(), "weights": w.clone()})
        z, udf = cat_z_vals(p, rays_o, rays_d, z, z_new, udf, last=(i + 1 == up_sample_steps))
    return z


@torch.no_grad()
def importance_sample_mix(p: UDFParams, s: "ScalarParams", rays_o, rays_d, z, sample_dist,
                          n_importance, up_sample_steps, sdf2alpha_type: str = "numerical"):
    """upsampling_type="mix" (udf_renderer_blending.py:843-918): S occlusion-UNaware steps (density
    peaks everywhere udf ~ 0) followed by one occlusion-aware step; k = n_importance // (S+1)."""
    B, n0 = z.shape
    S = up_sample_steps
    k = n_importance // (S + 1)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]
    udf = udf_forward(p, pts.reshape(-1, 3))[0][:, 0].reshape(B, n0)
    gamma = s.gamma_val()
    for i in range(S):
        z_new = up_sample_no_occ_aware(rays_o, rays_d, z, udf, sample_dist, k, 64 * 2 ** i,
                                       64 * 2 ** (i + 1), gamma)
        z, udf = cat_z_vals(p, rays_o, rays_d, z, z_new, udf, last=False)
    for i in range(S - 1, S):
        z_new = up_sample_unbias(rays_o, rays_d, z, udf, sample_dist, k, 64 * 2 ** i, 64 * 2 ** (i + 1),
                                 20 if i < 4 else 10, sdf2alpha_type=sdf2alpha_type)
        z, udf = cat_z_vals(p, rays_o, rays_d, z, z_new, udf, last=(i + 1 == S))
    return z


# --------------------------------------------------------------------------- #
# a13  render_core                 udf_renderer_blending.py:418-677
# --------------------------------------------------------------------------- #
def render_core(p: UDFParams, s: ScalarParams, rays_o, rays_d, z, sample_dist,
                cos_anneal_ratio=None, flip_saturation=0.0, near_surface=0.05,
                sparse_scale_factor=25000.0, use_unbias_render=True,
                use_norm_grad_for_cosine=False, sdf2alpha_type="numerical",
                udf_and_grad=None) -> Dict[str, torch.Tensor]:
    """``udf_and_grad`` (optional) injects (udf[P,1], grad[P,3]) so that the post-MLP
    stage can be checked in isolation against a CUDA kernel fed identical inputs."""
    B, n = z.shape
    last = torch.full_like(z[:, :1], float(sample_dist))
    dists = torch.cat([z[:, 1:] - z[:, :-1], last], -1)                    # :435-444
    mid_z = z + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)

    if udf_and_grad is None:
        udf = udf_forward(p, pts)[0][:, :1]                                # :457
        grad = udf_gradient(p, pts)                                        # :461
    else:
        udf, grad = udf_and_grad
    grad_mag = torch.linalg.norm(grad, ord=2, dim=-1, keepdim=True)
    grad_norm = grad / (grad_mag + 1e-5)

    inv_s = s.inv_s().reshape(1, 1)
    beta = s.beta_val()
    gamma = s.gamma_val()
    alpha2 = sdf2alpha_numerical if sdf2alpha_type == "numerical" else sdf2alpha_theorical

    flip_sign = None
    if use_unbias_render:
        cos_src = grad_norm if use_norm_grad_for_cosine else grad
        true_cos = (dirs * cos_src).sum(-1, keepdim=True)                  # :479-482
        with torch.no_grad():
            c = (dirs * grad_norm).sum(-1, keepdim=True)
            flip_sign = torch.sign(c) * -1
            flip_sign[flip_sign == 0] = 1                                  # :484-489
        raw_occ = udf2logistic(udf, beta, 1.0, 1.0).reshape(B, n)          # :492
        alpha_occ = 1.0 - torch.exp(-F.relu(raw_occ) * gamma * dists)      # :497
        vis_mask = (true_cos < 0.01).to(z.dtype).reshape(B, n)             # :500-504
        vis_mask = torch.cat([vis_mask[:, 1:], torch.ones_like(vis_mask[:, :1])], -1)
        vis_prob = _excl_cumprod((1.0 - alpha_occ + flip_saturation * vis_mask).clip(0, 1) + 1e-7)
        vis_prob = vis_prob.clip(0, 1)                                     # :511-528
        neg_abs_cos = -1 * torch.abs(true_cos)
        a_plus = alpha2(udf, neg_abs_cos, dists.reshape(-1, 1), inv_s, cos_anneal_ratio).reshape(B, n)
        a_minus = alpha2(-udf, neg_abs_cos, dists.reshape(-1, 1), inv_s, cos_anneal_ratio).reshape(B, n)
        alpha = a_plus * vis_prob + a_minus * (1 - vis_prob)               # :545
        udf = udf.reshape(B, n)
    else:
        udf = udf.reshape(B, n)
        raw_occ = udf2logistic(udf, beta, 1.0, 1.0).reshape(B, n)
        alpha = 1.0 - torch.exp(-F.relu(raw_occ) * gamma * dists)          # :551-559

    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1).reshape(B, n)
    inside_sphere = (pts_norm < 2.0).to(z.dtype)                           # :568
    relax_inside = (pts_norm < 2.4).to(z.dtype)                            # :569
    near = (udf < near_surface).to(z.dtype).detach()                       # :570

    weights = alpha * _excl_cumprod(1.0 - alpha + 1e-7)                    # :593-602
    edge = weights.sum(dim=-1, keepdim=True)     # sampled_edge == 1  (:561,:606)
    depth = (mid_z * weights).sum(dim=1, keepdim=True)                     # :607

    g3 = grad.reshape(B, n, 3)
    gerr = (torch.linalg.norm(g3, ord=2, dim=-1) - 1.0) ** 2               # :612-617
    gradient_error = (relax_inside * gerr).sum() / (relax_inside.sum() + 1e-5)
    gradient_error_ns = (near * gerr).sum() / (near.sum() + 1e-5)          # :618-625
    g_flip = flip_sign.reshape(B, n, 1) * g3 if flip_sign is not None else g3
    sparse_error = torch.exp(-sparse_scale_factor * udf).sum(dim=1).mean()  # :642-644

    return {
        "udf": udf, "edge": edge, "weights": weights,
        "s_val": (1.0 / inv_s).expand(B * n, 1), "beta": 1.0 / beta, "gamma": gamma,
        "depth": depth, "gradient_error": gradient_error,
        "gradient_error_near_surface": gradient_error_ns,
        "normals": (g_flip * weights[:, :, None]).sum(dim=1),
        "gradients": g3, "gradients_flip": g_flip, "inside_sphere": inside_sphere,
        "gradient_mag": grad_mag.reshape(B, n), "alpha": alpha, "mid_z_vals": mid_z,
        "dists": dists, "sparse_error": sparse_error,
    }


# --------------------------------------------------------------------------- #
# a8  render                       udf_renderer_blending.py:679-800
# --------------------------------------------------------------------------- #
@dataclass
class RenderConfig:
    n_samples: int = 64
    n_importance: int = 50
    up_sample_steps: int = 5
    perturb: float = 1.0
    sdf2alpha_type: str = "numerical"
    upsampling_type: str = "classical"
    sparse_scale_factor: float = 25000.0
    use_norm_grad_for_cosine: bool = False
    use_unbias_render: bool = True
    near_surface: float = 0.05


def coarse_z(near: torch.Tensor, far: torch.Tensor, n_samples: int,
             t_rand: Optional[torch.Tensor]) -> torch.Tensor:
    """udf_renderer_blending.py:705-720.  ``t_rand`` = rand([B,1]) - 0.5 or None."""
    lin = torch.linspace(0.0, 1.0, n_samples, dtype=near.dtype).to(near.device)
    z = near + (far - near) * lin[None, :]
    if t_rand is not None:
        z = z + t_rand * 2.0 / n_samples
    return z


def render(p: UDFParams, s: ScalarParams, cfg: RenderConfig, rays_o, rays_d, near, far,
           depth_scale, cos_anneal_ratio=None, flip_saturation=0.0,
           t_rand: Optional[torch.Tensor] = None, trace: Optional[list] = None,
           z_override: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """near/far: [B,1] (or [1,1]) tensors.  ``t_rand``: the (rand-0.5) jitter the reference
    draws from the global CPU generator (:719) -- passed in so results are reproducible."""
    sample_dist = ((far - near) / cfg.n_samples).mean().item()             # :704
    z = coarse_z(near, far, cfg.n_samples, t_rand if cfg.perturb > 0 else None)
    if z_override is not None:
        z = z_override
    elif cfg.n_importance > 0:
        if cfg.upsampling_type == "mix":
            z = importance_sample_mix(p, s, rays_o, rays_d, z, sample_dist, cfg.n_importance,
                                      cfg.up_sample_steps, cfg.sdf2alpha_type)
        else:
            z = importance_sample(p, rays_o, rays_d, z, sample_dist, cfg.n_importance,
                                  cfg.up_sample_steps, cfg.use_unbias_render, trace=trace,
                                  sdf2alpha_type=cfg.sdf2alpha_type)
    r = render_core(p, s, rays_o, rays_d, z, sample_dist, cos_anneal_ratio, flip_saturation,
                    cfg.near_surface, cfg.sparse_scale_factor, cfg.use_unbias_render,
                    cfg.use_norm_grad_for_cosine, cfg.sdf2alpha_type)
    w = r["weights"]
    return {
        "udf": r["udf"], "edge": r["edge"],
        "weight_sum": w.sum(dim=-1, keepdim=True), "weight_sum_fg_bg": w.sum(dim=-1, keepdim=True),
        "depth": r["depth"] * depth_scale, "variance": r["s_val"], "beta": r["beta"],
        "gamma": r["gamma"], "normals": r["normals"], "gradients": r["gradients"],
        "gradients_flip": r["gradients_flip"], "weights": w,
        "gradient_error": r["gradient_error"],
        "gradient_error_near_surface": r["gradient_error_near_surface"],
        "inside_sphere": r["inside_sphere"], "gradient_mag": r["gradient_mag"],
        "mid_z_vals": r["mid_z_vals"], "dists": r["dists"],
        # not returned by the reference's render(); kept for stage-level checks
        "_alpha": r["alpha"], "_sparse_error": r["sparse_error"], "_z_vals": z,
    }


# --------------------------------------------------------------------------- #
# a14  RenderingNetwork (dead in the reference; standalone op)  udf_model.py:177-209
# --------------------------------------------------------------------------- #
def rendering_network_forward(W: Sequence[torch.Tensor], b: Sequence[torch.Tensor], mode: str,
                              points, normals, view_dirs, feats, multires_view=4,
                              squeeze_out=True, d_out=1):
    if multires_view > 0 and mode != "no_view_dir":
        view_dirs = posenc(view_dirs, multires_view)
    if mode == "idr":
        x = torch.cat([points, view_dirs, normals, -1 * normals, feats], -1)
    elif mode == "no_view_dir":
        x = torch.cat([points, normals, -1 * normals, feats], -1)
    else:  # no_normal
        x = torch.cat([points, view_dirs, feats], -1)
    for l in range(len(W)):
        x = F.linear(x, W[l], b[l])
        if l < len(W) - 1:
            x = F.relu(x)
    x = x[:, :d_out]
    return torch.sigmoid(x) if squeeze_out else x


# --------------------------------------------------------------------------- #
# synthetic inputs -- shared by tests, smoke and bench
# --------------------------------------------------------------------------- #
def synthetic_rays(B: int, seed: int = 1234, dtype=torch.float32):
    """Random cameras on a radius-2.5 sphere looking at the (jittered) origin."""
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(B, 3, generator=g, dtype=torch.float64)
    c = 2.5 * c / c.norm(dim=-1, keepdim=True)
    target = 0.3 * torch.randn(B, 3, generator=g, dtype=torch.float64)
    fwd = target - c
    fwd = fwd / fwd.norm(dim=-1, keepdim=True)
    # pinhole spread of +-20 degrees around the optical axis
    a = torch.randn(B, 3, generator=g, dtype=torch.float64)
    a = a - (a * fwd).sum(-1, keepdim=True) * fwd
    a = a / a.norm(dim=-1, keepdim=True)
    ang = (torch.rand(B, 1, generator=g, dtype=torch.float64) * 2 - 1) * math.radians(20.0)
    d = fwd * torch.cos(ang) + a * torch.sin(ang)
    d = d / d.norm(dim=-1, keepdim=True)
    return c.to(dtype).contiguous(), d.to(dtype).contiguous()


def synthetic_t_rand(B: int, seed: int = 7) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 1, generator=g) - 0.5


def perturbed_params(p: UDFParams, sigma: float = 0.02, seed: int = 1) -> UDFParams:
    """Second weight set of SURVEY §8d: geometric init + noise on every direction tensor so the
    high-frequency PE columns, the skip columns and the biases are exercised.  The noise on a
    PE column of frequency 2^j is scaled by 2^-j (a 1/f spectrum, as in a trained network) --
    otherwise |grad udf| reaches the hundreds and nothing resembles a distance field."""
    g = torch.Generator().manual_seed(seed)
    pe = p.v[0].shape[1]
    col_scale = torch.ones(pe)
    for j in range(p.multires):
        col_scale[3 + 6 * j: 9 + 6 * j] = 2.0 ** (-j)
    v, b = [], []
    for l, t in enumerate(p.v):
        noise = sigma * torch.randn(t.shape, generator=g)
        if l == 0:
            noise = noise * col_scale[None, :]
        elif l in p.skip_in:
            noise[:, -pe:] = noise[:, -pe:] * col_scale[None, :]
        v.append(t + noise)
    for t in p.b:
        b.append(t + sigma * torch.randn(t.shape, generator=g))
    return UDFParams(v, [t.clone() for t in p.g], b, p.multires, p.skip_in, p.scale, p.udf_type)
