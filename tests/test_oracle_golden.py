"""Pin oracle/emap_oracle.py against the fixtures generated FROM THE REFERENCE
(tests/golden/make_golden.py).  CPU only.

Bit-exact (torch.equal) wherever the oracle issues the same torch ops in the same order as the
reference; a tight 2e-6 bound elsewhere (the only such place is where the reference multiplies by
sampled_edge == 1 / reshapes through views, which cannot change values -- so in practice all of
these are exact, and the test says so)."""
import pytest
import torch

from oracle import emap_oracle as O
from tests.helpers import maxdiff, oracle_params, oracle_scalars

torch.set_num_threads(8)


def test_embedder(golden):
    g = golden("embed")
    assert torch.equal(O.posenc(g["x"], 10), g["pe10"])
    assert torch.equal(O.posenc(g["x"], 6), g["pe6"])


@pytest.mark.parametrize("tag,pert", [("init", False), ("pert", True)])
def test_mlp_forward_and_gradient(golden, tag, pert):
    g = golden(f"mlp_{tag}")
    p = oracle_params(pert)
    out, pe = O.udf_forward(p, g["x"])
    assert torch.equal(out, g["out"])
    assert torch.equal(pe, g["pe"])
    grad = O.udf_gradient(p, g["x"]).detach()
    assert torch.equal(grad, g["grad"][:, 0])


@pytest.mark.parametrize("tag,pert", [("init", False), ("pert", True)])
def test_mlp_double_backward(golden, tag, pert):
    g = golden(f"mlp_{tag}")
    p = oracle_params(pert).requires_grad_(True)
    y = O.udf_forward(p, g["x"])[0]
    gg = O.udf_gradient(p, g["x"])
    loss = (g["cu"] * y).sum() + (g["cg"] * gg).sum()
    assert torch.equal(loss.detach(), g["loss"])
    grads = torch.autograd.grad(loss, p.tensors())
    names = []
    for l in range(p.n_linear):
        names += [f"lin{l}.bias", f"lin{l}.parametrizations.weight.original0",
                  f"lin{l}.parametrizations.weight.original1"]
    for n, gr in zip(names, grads):
        ref = g[f"dgrad.{n}"]
        assert maxdiff(gr, ref) <= 2e-6 * (1 + float(ref.abs().max())), n


def test_mlp_multires6(golden):
    g = golden("mlp_mr6_pert")
    p = oracle_params(True, multires=6)
    assert torch.equal(O.udf_forward(p, g["x"])[0], g["out"])
    assert torch.equal(O.udf_gradient(p, g["x"]).detach(), g["grad"][:, 0])


@pytest.mark.parametrize("k", [10, 16, 32])
def test_sample_pdf(golden, k):
    g = golden(f"sample_pdf_k{k}")
    s, inds = O.sample_pdf_det(g["bins"], g["weights"], k, return_inds=True)
    assert torch.equal(inds, g["inds"])
    assert torch.equal(s, g["samples"])


@pytest.mark.parametrize("tag,pert,n0,ni,steps", [("init_64_50_5", False, 64, 50, 5),
                                                   ("pert_64_64_4", True, 64, 64, 4),
                                                   ("pert_128_128_4", True, 128, 128, 4)])
def test_upsampling_trace(golden, tag, pert, n0, ni, steps):
    g = golden(f"upsample_{tag}")
    p = oracle_params(pert)
    o, d = g["rays_o"], g["rays_d"]
    z = O.coarse_z(g["near"], g["far"], n0, g["t_rand"])
    assert torch.equal(z, g["z0"])
    sd = float(g["sample_dist"])
    trace = []
    zf = O.importance_sample(p, o, d, z, sd, ni, steps, trace=trace)
    assert torch.equal(trace[0]["udf"], g["udf0"])
    for i in range(steps):
        assert torch.equal(trace[i]["z_new"], g[f"z_new{i}"]), i
        assert torch.equal(trace[i]["z"], g[f"z{i}"]) if i > 0 else True
        if 0 < i:
            assert torch.equal(trace[i]["udf"], g[f"udf{i}"]), i
    assert torch.equal(zf, g["z_final"])
    # stage in isolation: one step from the reference's own (z, udf)
    for i in range(steps):
        zi = g["z0"] if i == 0 else g[f"z{i}"]
        ui = g["udf0"] if i == 0 else g[f"udf{i}"]
        inv_s, beta, gamma = O.upsample_schedule(i, steps)
        zn = O.up_sample_unbias(o, d, zi, ui, sd, ni // steps, inv_s, beta, gamma)
        assert torch.equal(zn, g[f"z_new{i}"])


RENDER_CASES = [
    ("init_64_50_5", False, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5)),
    ("pert_64_64_4", True, 10, dict(n_samples=64, n_importance=64, up_sample_steps=4)),
    ("pert_64_0", True, 10, dict(n_samples=64, n_importance=0, up_sample_steps=5)),
    ("pert_128_128_4", True, 10, dict(n_samples=128, n_importance=128, up_sample_steps=4)),
    ("mr6_64_50_5", True, 6, dict(n_samples=64, n_importance=50, up_sample_steps=5)),
    ("var_biased", True, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5,
                                  use_unbias_render=False)),
    ("var_theorical", True, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5,
                                     sdf2alpha_type="theorical")),
    ("var_normgrad", True, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5,
                                    use_norm_grad_for_cosine=True)),
    ("var_mix", True, 10, dict(n_samples=64, n_importance=60, up_sample_steps=5, upsampling_type="mix")),
]

OUT_KEYS = ["udf", "edge", "weight_sum", "weight_sum_fg_bg", "depth", "beta", "gamma", "normals",
            "gradients", "gradients_flip", "weights", "gradient_error",
            "gradient_error_near_surface", "inside_sphere", "gradient_mag", "mid_z_vals", "dists"]


def run_oracle_render(g, pert, multires, cfgkw, requires_grad=False):
    p = oracle_params(pert, multires)
    s = oracle_scalars()
    if requires_grad:
        p.requires_grad_(True)
        for t in (s.variance, s.beta, s.gamma):
            t.requires_grad_(True)
    cfg = O.RenderConfig(**cfgkw)
    car = float(g["cos_anneal_ratio"])
    out = O.render(p, s, cfg, g["rays_o"], g["rays_d"], g["near"], g["far"], g["depth_scale"],
                   cos_anneal_ratio=None if car < 0 else car,
                   flip_saturation=float(g["flip_saturation"]), t_rand=g["t_rand"])
    return p, s, out


@pytest.mark.parametrize("tag,pert,multires,cfgkw", RENDER_CASES)
def test_render_outputs(golden, tag, pert, multires, cfgkw):
    g = golden(f"render_{tag}")
    _, _, out = run_oracle_render(g, pert, multires, cfgkw)
    for k in OUT_KEYS:
        ref = g[f"out.{k}"]
        got = out[k].detach()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert torch.equal(got, ref), (k, maxdiff(got, ref))
    assert torch.equal(out["variance"][:1].detach(), g["out.variance0"])
    B, n = g["out.udf"].shape
    assert out["variance"].shape == (B * n, 1)


@pytest.mark.parametrize("tag,pert,multires,cfgkw", RENDER_CASES[:3])
def test_render_param_grads(golden, tag, pert, multires, cfgkw):
    g = golden(f"render_{tag}")
    p, s, out = run_oracle_render(g, pert, multires, cfgkw, requires_grad=True)
    B = g["rays_o"].shape[0]
    loss = (torch.nn.functional.mse_loss(out["edge"], g["true_edge"])
            + 0.01 * out["gradient_error_near_surface"] + 0.1 * out["gradient_error"]
            + 0.05 * (out["depth"] * torch.linspace(-1, 1, B).reshape(B, 1)).mean()
            + 0.05 * (out["normals"] * g["rays_o"]).mean())
    assert torch.equal(loss.detach(), g["loss"])
    tensors = p.tensors() + [s.variance, s.beta, s.gamma]
    grads = torch.autograd.grad(loss, tensors)
    names = []
    for l in range(p.n_linear):
        names += [f"lin{l}.bias", f"lin{l}.parametrizations.weight.original0",
                  f"lin{l}.parametrizations.weight.original1"]
    names += ["variance", "beta", "gamma"]
    for n, gr in zip(names, grads):
        ref = g[f"dloss.{n}"]
        assert maxdiff(gr, ref) <= 2e-6 * (1 + float(ref.abs().max())), n


def test_rendering_network(golden):
    g = golden("rendering_network")
    W, b = [], []
    for l in range(5):
        W.append(torch._weight_norm(g[f"sd.lin{l}.parametrizations.weight.original1"],
                                    g[f"sd.lin{l}.parametrizations.weight.original0"], 0))
        b.append(g[f"sd.lin{l}.bias"])
    c = O.rendering_network_forward(W, b, "no_normal", g["pts"], g["normals"], g["view_dirs"],
                                    g["feat"])
    assert torch.equal(c, g["color"])


def test_scalars(golden):
    g = golden("scalars")
    s = oracle_scalars()
    assert torch.equal(s.inv_s().reshape(1, 1), g["inv_s"].clip(1e-6, 1e6))
    assert torch.equal(s.beta_val(), g["beta"].clip(1e-6, 1e6))
    assert torch.equal(s.gamma_val(), g["gamma"].clip(1e-6, 1e6))


# ----------------------------------------------------------------------------- SURVEY §8f rows 1-2
def test_extract_grid_oracle_matches_reference(golden):
    """oracle/extract_oracle.py reproduces the reference's get_udf_normals_grid output (same seed ->
    same randn offsets): df / coordinates / -sign(grad) "normals" bit-exactly, line directions up to the
    SVD sign."""
    from oracle import extract_oracle as E
    from tests.helpers import oracle_params
    g = golden("extract_grid")
    p = oracle_params(True)
    func = lambda x: (O.udf_forward(p, x)[0][:, :1], None, None)   # noqa: E731

    def func_grad(xyz):
        gr = O.udf_gradient(p, xyz).detach().reshape(-1, 1, 3)       # UDFNetwork.gradient -> [P,1,3]
        return gr / (torch.linalg.norm(gr, ord=2, dim=-1, keepdim=True) + 1e-5)

    N = int(g["N"])
    df, ld, vecs, samples, vs = E.udf_normals_grid(func, func_grad, N, float(g["udf_threshold"]), True,
                                                   int(g["sampling_N"]), float(g["sampling_delta"]),
                                                   offsets=g["offsets"])
    assert torch.equal(samples[:, :3], g["samples"][:, :3])
    assert torch.equal(df, g["df_values"])
    assert torch.equal(vecs, g["vecs"])
    a, b = ld.reshape(-1, 3), g["line_directions"].reshape(-1, 3)
    assert float(torch.minimum((a - b).abs().amax(1), (a + b).abs().amax(1)).max()) <= 1e-4
    assert float(vs) == float(g["voxel_size"])


def test_raygen_oracle_matches_reference(golden):
    from oracle import extract_oracle as E
    g = golden("raygen")
    i = int(g["img_idx"])
    out = E.rays_from_pixels(g["pixels_x"], g["pixels_y"], g["edges"][i], g["intrinsics_inv"][i], g["pose"][i],
                             int(g["H"]), int(g["W"]))
    for k in ("rays_o", "rays_v", "edge", "rays_ndc_uv", "rays_norm_XYZ_cam", "depth_scale"):
        assert torch.equal(out[k], g[k]), k
