#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (``/root/reference`` is read-only there and does not
exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules (``src.models.udf_model``,
``src.models.udf_renderer_blending`` -- they need only torch + numpy), runs them on the
CPU in fp32 on small seeded inputs, and stores inputs + outputs as ``.npz``.  Nothing from
the reference is copied into the repo; only its numerical outputs are.

The reference ships no tests or golden vectors of its own, so these files are
what pins ``oracle/emap_oracle.py`` (tests/test_oracle_golden.py) and, through it and
directly, the CUDA kernels (tests/test_gpu_*.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from src.models.udf_model import (BetaNetwork, RenderingNetwork, SingleVarianceNetwork,  # noqa: E402
                                  UDFNetwork)
from src.models.udf_renderer_blending import UDFRendererBlending, sample_pdf  # noqa: E402
from src.models.embedder import get_embedder  # noqa: E402

from oracle import emap_oracle as O  # noqa: E402  (only for the synthetic input generators)

torch.set_default_dtype(torch.float32)
torch.set_num_threads(8)

NET_KW = dict(d_in=3, d_out=1, d_hidden=256, n_layers=8, skip_in=[4], multires=10, bias=0.5,
              scale=1.0, geometric_init=True, weight_norm=True, udf_type="abs")


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


ONLY = os.environ.get("EMAP_GOLDEN_ONLY")   # comma-separated fixture names; default: all


def save(name, **arrs):
    if ONLY and name not in ONLY.split(","):
        return
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **npify(arrs))
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.0f} KiB  keys={len(arrs)}")


def build_net(multires=10, perturbed=False):
    torch.manual_seed(0)
    kw = dict(NET_KW)
    kw["multires"] = multires
    net = UDFNetwork(**kw)
    if perturbed:
        p = O.UDFParams.from_state_dict(net.state_dict(), multires=multires)
        p2 = O.perturbed_params(p)
        sd = net.state_dict()
        for l in range(p2.n_linear):
            sd[f"lin{l}.parametrizations.weight.original1"] = p2.v[l]
            sd[f"lin{l}.parametrizations.weight.original0"] = p2.g[l]
            sd[f"lin{l}.bias"] = p2.b[l]
        net.load_state_dict(sd)
    return net


def scalar_nets():
    var = SingleVarianceNetwork(0.3)
    beta = BetaNetwork(0.5, 0.3, 0.3, 5e-5, True, True, False)
    return var, beta


def rays(B, far=6.0):
    o, d = O.synthetic_rays(B)
    near = torch.full((B, 1), 0.05)
    farr = torch.full((B, 1), far)
    return o, d, near, farr


def main():
    # ------------------------------------------------------------------ network weights
    net0 = build_net()
    save("net_init_state", **{k: v for k, v in net0.state_dict().items()})
    net6 = build_net(multires=6)
    save("net_init_state_mr6", **{k: v for k, v in net6.state_dict().items()})

    # ------------------------------------------------------------------ a1 embedder
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(96, 3, generator=g) * 2 - 1) * 3.0
    emb10, d10 = get_embedder(10, 3)
    emb6, d6 = get_embedder(6, 3)
    save("embed", x=x, pe10=emb10(x), pe6=emb6(x))

    # ------------------------------------------------------------------ a3-a5 MLP fwd / gradient
    for tag, pert in (("init", False), ("pert", True)):
        net = build_net(perturbed=pert)
        g = torch.Generator().manual_seed(12)
        x = (torch.rand(384, 3, generator=g) * 2 - 1) * 1.5
        out, pe = net(x)
        udf, feat, _ = net.udf(x)
        grad = net.gradient(x.clone()).detach()
        # parameter gradients of a scalar that depends on both udf and grad (double backward)
        cu = torch.randn(384, 1, generator=g)
        cg = torch.randn(384, 3, generator=g)
        xx = x.clone()
        net.zero_grad()
        y = net(xx)[0]
        gg = net.gradient(xx).squeeze(1)
        loss = (cu * y).sum() + (cg * gg).sum()
        loss.backward()
        pg = {f"dgrad.{k}": v.grad for k, v in net.named_parameters()}
        save(f"mlp_{tag}", x=x, out=out, pe=pe, udf=udf, grad=grad, cu=cu, cg=cg, loss=loss.detach(),
             **pg)
    net = build_net(multires=6, perturbed=True)
    g = torch.Generator().manual_seed(13)
    x = (torch.rand(128, 3, generator=g) * 2 - 1) * 1.5
    save("mlp_mr6_pert", x=x, out=net(x)[0], grad=net.gradient(x.clone()).detach())

    # ------------------------------------------------------------------ a11b sample_pdf
    g = torch.Generator().manual_seed(14)
    B, n = 48, 64
    bins = torch.sort(torch.rand(B, n, generator=g) * 6, dim=-1)[0]
    w = torch.rand(B, n - 1, generator=g) ** 8          # peaky
    w[:4] = 0.0                                         # all-zero rows -> uniform pdf
    w[4:8, :] = 0.0
    w[4:8, 17] = 1.0                                    # single spike
    for k in (10, 16, 32):
        # recompute inds the way the reference does, to store them (sample_pdf returns samples only)
        ww = w + 1e-5
        pdf = ww / torch.sum(ww, -1, keepdim=True)
        cdf = torch.cat([torch.zeros(B, 1), torch.cumsum(pdf, -1)], -1)
        u = torch.linspace(0.5 / k, 1 - 0.5 / k, steps=k).expand(B, k).contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        save(f"sample_pdf_k{k}", bins=bins, weights=w, samples=sample_pdf(bins, w, k, det=True),
             inds=inds)

    # ------------------------------------------------------------------ a10/a12/a9 up-sampling
    var, beta = scalar_nets()
    for tag, pert, n0, ni, steps, B in (("init_64_50_5", False, 64, 50, 5, 24),
                                        ("pert_64_64_4", True, 64, 64, 4, 24),
                                        ("pert_128_128_4", True, 128, 128, 4, 8)):
        net = build_net(perturbed=pert)
        r = UDFRendererBlending(None, net, var, beta, n_samples=n0, n_importance=ni, n_outside=0,
                                up_sample_steps=steps, perturb=1.0, device="cpu")
        o, d, near, far = rays(B)
        t_rand = O.synthetic_t_rand(B)
        sample_dist = ((far - near) / n0).mean().item()
        z = near + (far - near) * torch.linspace(0, 1, n0)[None, :] + t_rand * 2.0 / n0
        with torch.no_grad():
            pts = o[:, None, :] + d[:, None, :] * z[..., :, None]
            udf = net(pts.reshape(-1, 3))[0].reshape(B, n0)
            out = {"rays_o": o, "rays_d": d, "near": near, "far": far, "t_rand": t_rand,
                   "z0": z, "udf0": udf, "sample_dist": sample_dist}
            zc, uc = z, udf
            k = ni // steps
            for i in range(steps):
                inv_s, bet, gam = 64 * 2 ** i, 64 * 2 ** (i + 1), float(np.clip(20 * 2 ** (steps - i), 20, 320))
                zn = r.up_sample_unbias(o, d, zc, uc, sample_dist, k, inv_s, bet, gam)
                out[f"z_new{i}"] = zn
                zc, uc = r.cat_z_vals(o, d, zc, zn, uc, last=(i + 1 == steps))
                out[f"z{i + 1}"] = zc
                if i + 1 < steps:
                    out[f"udf{i + 1}"] = uc
            zf = r.importance_sample(o, d, z, sample_dist)
            assert torch.equal(zf, zc)
            out["z_final"] = zf
        save(f"upsample_{tag}", **out)

    # ------------------------------------------------------------------ a13/a8 render_core / render
    def run_render(tag, pert, n0, ni, steps, B, flip, cos_ratio, multires=10, far_v=6.0, grads=True,
                   **rkw):
        net = build_net(multires=multires, perturbed=pert)
        var, beta = scalar_nets()
        r = UDFRendererBlending(None, net, var, beta, n_samples=n0, n_importance=ni, n_outside=0,
                                up_sample_steps=steps, perturb=1.0, device="cpu", **rkw)
        o, d, near, far = rays(B, far_v)
        depth_scale = torch.linspace(0.5, 1.5, B).reshape(B, 1)
        t_rand = O.synthetic_t_rand(B)
        torch.manual_seed(7)           # reference draws rand([B,1]) from the global CPU generator
        t_chk = torch.rand(B, 1) - 0.5
        assert torch.equal(t_chk, t_rand)
        torch.manual_seed(7)
        out = r.render(o, d, near, far, depth_scale, cos_anneal_ratio=cos_ratio, flip_saturation=flip)
        fx = {"rays_o": o, "rays_d": d, "near": near, "far": far, "t_rand": t_rand,
              "depth_scale": depth_scale, "flip_saturation": flip,
              "cos_anneal_ratio": -1.0 if cos_ratio is None else cos_ratio}
        fx.update({f"out.{k}": v for k, v in out.items() if k != "variance"})
        fx["out.variance0"] = out["variance"][:1]
        if grads:
            ge = torch.Generator().manual_seed(21)
            true_edge = torch.rand(B, 1, generator=ge)
            loss = (torch.nn.functional.mse_loss(out["edge"], true_edge)
                    + 0.01 * out["gradient_error_near_surface"] + 0.1 * out["gradient_error"]
                    + 0.05 * (out["depth"] * torch.linspace(-1, 1, B).reshape(B, 1)).mean()
                    + 0.05 * (out["normals"] * o).mean())
            for m in (net, var, beta):
                m.zero_grad()
            loss.backward()
            fx["true_edge"] = true_edge
            fx["loss"] = loss.detach()
            for k, v in net.named_parameters():
                fx[f"dloss.{k}"] = v.grad
            fx["dloss.variance"] = var.variance.grad
            fx["dloss.beta"] = beta.beta.grad
            fx["dloss.gamma"] = beta.gamma.grad
        save(f"render_{tag}", **fx)

    run_render("init_64_50_5", False, 64, 50, 5, 16, 0.0, 1.0)
    run_render("pert_64_64_4", True, 64, 64, 4, 16, 0.9, 0.6)
    run_render("pert_64_0", True, 64, 0, 5, 16, 0.9, None)
    run_render("pert_128_128_4", True, 128, 128, 4, 6, 1.0, 1.0)      # the bench configuration (BASELINE configs[3])
    # Replica-style: multires=6, far=2.5
    run_render("mr6_64_50_5", True, 64, 50, 5, 8, 0.9, 1.0, multires=6, far_v=2.5, grads=False)
    # off-default variants
    run_render("var_biased", True, 64, 50, 5, 8, 0.9, 1.0, grads=False, use_unbias_render=False)
    run_render("var_theorical", True, 64, 50, 5, 8, 0.9, 0.7, grads=False, sdf2alpha_type="theorical")
    run_render("var_normgrad", True, 64, 50, 5, 8, 0.9, 1.0, grads=False, use_norm_grad_for_cosine=True)
    # upsampling_type="mix": n_importance // (steps+1) = 10 per step, 6 steps -> 64+60 samples, while the
    # reference's render() still reports n_samples + n_importance = 114 in weight_sum (quirk kept)
    run_render("var_mix", True, 64, 60, 5, 8, 0.9, 1.0, grads=False, upsampling_type="mix")

    # ------------------------------------------------------------------ a14 RenderingNetwork
    torch.manual_seed(3)
    rn = RenderingNetwork(d_feature=256, mode="no_normal", d_in=6, d_out=1, d_hidden=128, n_layers=4,
                          weight_norm=True, multires_view=4, squeeze_out=True)
    g = torch.Generator().manual_seed(15)
    P = 64
    pts, nrm, vd = (torch.randn(P, 3, generator=g) for _ in range(3))
    feat = torch.randn(P, 256, generator=g)
    save("rendering_network", pts=pts, normals=nrm, view_dirs=vd, feat=feat,
         color=rn(pts, nrm, vd, feat), **{f"sd.{k}": v for k, v in rn.state_dict().items()})

    # ------------------------------------------------------------------ §8f row 1: grid query of extract_edge
    # the reference's own get_udf_normals_grid (loaded from its source file: the package __init__ pulls
    # open3d), driven by the reference UDFNetwork exactly as runner_udf.py:520-526 does, on the CPU
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_extract_pointcloud", "/root/reference/src/edge_extraction/extract_pointcloud.py")
    ref_ep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_ep)
    net = build_net(perturbed=True)
    func = net.udf

    def func_grad(xyz):
        gradients = net.gradient(xyz)
        mag = torch.linalg.norm(gradients, ord=2, dim=-1, keepdim=True)
        return gradients / (mag + 1e-5)

    Ng, thr, SN = 14, 0.3, 50
    torch.manual_seed(31)
    df, ld, vecs, samples, vs = ref_ep.get_udf_normals_grid(func, func_grad, Ng, thr, is_linedirection=True,
                                                            sampling_N=SN, sampling_delta=0.005,
                                                            max_batch=4096, device="cpu")
    M = int((samples[:, 3] < thr).sum())
    torch.manual_seed(31)
    offsets = torch.randn((M, SN, 3))          # the draw the reference made (M <= max_batch: one batch)
    assert M <= 4096
    save("extract_grid", N=Ng, udf_threshold=thr, sampling_N=SN, sampling_delta=0.005, df_values=df,
         line_directions=ld, vecs=vecs, samples=samples.detach(), voxel_size=vs, offsets=offsets)

    # ------------------------------------------------------------------ §8f row 2: ray sampler (deterministic half)
    # gen_random_rays_patches_at is exec'd from the reference's dataset.py (the module itself imports cv2
    # etc.) with a stand-in `self`; pixels are recovered from the returned ndc coordinates
    import ast
    import random
    import types
    src = open("/root/reference/src/dataset/dataset.py").read()
    fn_node = None
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == "gen_random_rays_patches_at":
            fn_node = node
    mod = ast.Module(body=[fn_node], type_ignores=[])
    ns = {"torch": torch, "np": np, "random": random}
    exec(compile(mod, "dataset.py:gen_random_rays_patches_at", "exec"), ns)
    H, W, n_img = 48, 64, 3
    g = torch.Generator().manual_seed(41)
    edges = (torch.rand(n_img, H, W, 1, generator=g) > 0.9).float() * torch.rand(n_img, H, W, 1, generator=g)
    K = torch.eye(4).repeat(n_img, 1, 1)
    K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2] = 70.0, 72.0, W / 2 - 0.3, H / 2 + 0.2
    Kinv = torch.inverse(K)
    pose = torch.eye(4).repeat(n_img, 1, 1)
    for i in range(n_img):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        pose[i, :3, :3] = q
        pose[i, :3, 3] = torch.randn(3, generator=g)
    fake = types.SimpleNamespace(W=W, H=H, edges=edges, intrinsics_all=K, intrinsics_all_inv=Kinv,
                                 pose_all=pose, masks=None, device="cpu", image_pixels=H * W)
    torch.manual_seed(43)
    out = ns["gen_random_rays_patches_at"](fake, 1, 256, importance_sample=False)
    uv = out["rays_ndc_uv"]
    px = torch.round((uv[:, 0].double() + 1) / 2 * (W - 1)).long()
    py = torch.round((uv[:, 1].double() + 1) / 2 * (H - 1)).long()
    save("raygen", H=H, W=W, img_idx=1, edges=edges, intrinsics_inv=Kinv, pose=pose, pixels_x=px,
         pixels_y=py, rays_o=out["rays"]["rays_o"], rays_v=out["rays"]["rays_v"], edge=out["rays"]["edge"],
         rays_ndc_uv=uv, rays_norm_XYZ_cam=out["rays_norm_XYZ_cam"], depth_scale=out["depth_scale"])

    # ------------------------------------------------------------------ scalar nets
    var, beta = scalar_nets()
    save("scalars", inv_s=var(torch.zeros(1, 3)), beta=beta.get_beta(), gamma=beta.get_gamma(),
         variance=var.variance, beta_raw=beta.beta, gamma_raw=beta.gamma)


if __name__ == "__main__":
    main()
