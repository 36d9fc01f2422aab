"""GPU parity, round 2: the gaps the round-1 review listed, each bound = what tools/gpu/r2_parity_probe.py measured
on B200 (profiles/r02_parity.json) with the margin stated next to it.

  * per-sample weights / udf / gradients at the REFERENCE's sample positions for all nine render() cases
    (round 1 compared them only in the flat case), and end-to-end bounds that follow the measurements;
  * index flips of the isolated up-sampling steps are proven to be knife edges (a cdf entry within 2e-6 of the
    quantile), not just counted;
  * parameter gradients at the bench configuration (128+128/4) against a reference-generated fixture;
  * the eikonal-only gradient at PRODUCTION scale (1024 rays, loss weight 0.01: cotangents ~1e-8) against the
    fp64 oracle -- the case the unscaled fp16 backward of round 1 flushed to zero;
  * the bf16 network (BASELINE config C2): forward bounds and gradients;
  * the device status word that replaces the reference's pdb NaN guards;
  * row f4: a stand-in of the reference's validate() loop and checkpoint save/load (runner_udf.py:252-408).
"""
import os

import pytest
import torch

from tests.helpers import maxdiff
from tests.test_gpu_render import CASES, build, run_render

pytestmark = pytest.mark.gpu
dev = "cuda"


def l2rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def _render_core_at_reference_z(g, net, var, beta, r):
    z_ref = (g["out.mid_z_vals"] - 0.5 * g["out.dists"]).to(dev)
    sd = torch.tensor([float(g["out.dists"][0, -1])], device=dev)
    car = float(g["cos_anneal_ratio"])
    return r.render_core(g["rays_o"].to(dev), g["rays_d"].to(dev), z_ref, sd, net, var, beta_network=beta,
                         cos_anneal_ratio=None if car < 0 else car, flip_saturation=float(g["flip_saturation"]))


@pytest.mark.parametrize("tag,pert,multires,rkw", CASES)
def test_render_core_per_sample_at_reference_positions(golden, tag, pert, multires, rkw):
    """render_core (MLP + gradient + compositing) fed the reference's own sample positions: every per-sample
    tensor is comparable one to one -- for the hierarchical cases too."""
    g = golden(f"render_{tag}")
    net, var, beta, r = build(multires, pert, **rkw)
    with torch.no_grad():
        c = _render_core_at_reference_z(g, net, var, beta, r)
    c = {k: v.detach().cpu() for k, v in c.items()}
    assert maxdiff(c["mid_z_vals"], g["out.mid_z_vals"]) <= 2e-6
    assert maxdiff(c["dists"], g["out.dists"]) <= 2e-6
    # measured worst case over the nine fixtures (profiles/r02_parity.json) in brackets; bound = 2x
    assert maxdiff(c["udf"], g["out.udf"]) <= 6e-5                       # [2.7e-5]
    assert maxdiff(c["gradients"], g["out.gradients"]) <= 8e-5           # [3.7e-5]
    assert maxdiff(c["gradient_mag"], g["out.gradient_mag"]) <= 8e-5     # [3.2e-5]
    assert maxdiff(c["weights"], g["out.weights"]) <= 1.2e-5             # [5.6e-6]
    assert maxdiff(c["gradients_flip"], g["out.gradients_flip"]) <= 8e-5  # [3.7e-5]
    assert maxdiff(c["edge"], g["out.edge"]) <= 2e-4                     # [9.0e-5] (sum of up to 256 weights)
    assert maxdiff(c["normals"], g["out.normals"]) <= 1.5e-4             # [6.8e-5]
    assert torch.equal(c["inside_sphere"], g["out.inside_sphere"])
    ge = float(g["out.gradient_error"])
    assert abs(float(c["gradient_error"]) - ge) <= 5e-4 * max(1.0, ge)


@pytest.mark.parametrize("tag,pert,multires,rkw", CASES)
def test_render_end_to_end_measured_bounds(golden, tag, pert, multires, rkw):
    """render() end to end (own up-sampling).  Bounds = 2x the worst measured case (profiles/r02_parity.json):
    edge / weight_sum 8.6e-5, depth 4.9e-4 everywhere.  Sample positions: 7.4e-4 on the eight perturbed-network
    fixtures; 9.4e-3 on the pristine geometric initialisation (init_64_50_5: an exact sphere, whose rays carry
    long stretches of ~1e-5 weight where the inverse CDF is flat and a rounding-level change of the cdf moves a
    sample along the ray without changing any rendered quantity -- edge still agrees to 8.6e-5 there)."""
    g = golden(f"render_{tag}")
    net, var, beta, r = build(multires, pert, **rkw)
    with torch.no_grad():
        o = run_render(g, r)
    o = {k: v.detach().cpu() for k, v in o.items()}
    flat = rkw["n_importance"] == 0
    dz = (o["mid_z_vals"] - g["out.mid_z_vals"]).abs()
    assert float(dz.max()) <= (2e-6 if flat else (2e-2 if not pert else 1.5e-3))
    assert maxdiff(o["edge"], g["out.edge"]) <= 2e-4
    assert maxdiff(o["weight_sum"], g["out.weight_sum"]) <= 2e-4
    assert maxdiff(o["depth"], g["out.depth"]) <= 1e-3


@pytest.mark.parametrize("tag,n0,ni,steps", [("init_64_50_5", 64, 50, 5), ("pert_64_64_4", 64, 64, 4),
                                             ("pert_128_128_4", 128, 128, 4)])
def test_upsample_index_flips_are_knife_edges(golden, tag, n0, ni, steps):
    """each up-sampling step fed the reference's own (z, udf): an index may differ from the reference's only
    where a cdf entry sits within 2e-6 of the quantile (searchsorted on a knife edge: libm ulps decide).
    Measured on B200: ZERO flips on all three fixtures (13 steps, 18,624 look-ups) -- asserted exactly; the
    knife-edge proof stays in place for other hardware / libm versions.  New samples: <= 1.5e-6 from the
    reference's on the perturbed networks; on the pristine sphere initialisation (flat ~1e-5 weight stretches,
    see test_render_end_to_end_measured_bounds) 1.3e-3 in step 0 and 2.6e-5 in step 2."""
    from emap_b200 import ops
    from oracle import emap_oracle as O
    g = golden(f"upsample_{tag}")
    o, d = g["rays_o"].to(dev), g["rays_d"].to(dev)
    sd = torch.tensor([float(g["sample_dist"])], device=dev)
    k = ni // steps
    u = torch.linspace(0.5 / k, 1 - 0.5 / k, steps=k)
    flips = 0
    for i in range(steps):
        zi = g["z0"] if i == 0 else g[f"z{i}"]
        ui = g["udf0"] if i == 0 else g[f"udf{i}"]
        inv_s, bet, gam = O.upsample_schedule(i, steps)
        _, _, z_new, inds, w = ops.upsample_step(o, d, zi.to(dev), ui.to(dev), None, None, u.to(dev), k, sd,
                                                 inv_s, bet, gam, want_inds=True, want_weights=True)
        zr, ir, wr = O.up_sample_unbias(g["rays_o"], g["rays_d"], zi, ui, float(g["sample_dist"]), k,
                                        inv_s, bet, gam, return_aux=True)
        ww = wr.double() + 1e-5
        cdf = torch.cat([torch.zeros(ww.shape[0], 1, dtype=torch.float64),
                         torch.cumsum(ww / ww.sum(-1, keepdim=True), -1)], -1)
        for ray, j in (inds.cpu() != ir).nonzero().tolist():
            a, b = int(inds[ray, j]), int(ir[ray, j])
            assert abs(a - b) == 1, (i, ray, j, a, b)
            edge = float((cdf[ray, min(a, b)] - float(u[j])).abs())
            assert edge <= 2e-6, (i, ray, j, edge)
            flips += 1
        dz = maxdiff(z_new.cpu(), torch.sort(g[f"z_new{i}"], -1)[0])
        assert dz <= (3e-3 if tag.startswith("init") else 5e-6), (i, dz)
        assert maxdiff(w.cpu(), wr) <= 2e-6 * max(1e-3, float(wr.abs().max())), i      # measured 5.4e-7
    assert flips == 0, flips


NAMES = []
for _l in range(9):
    NAMES += [f"lin{_l}.bias", f"lin{_l}.parametrizations.weight.original0",
              f"lin{_l}.parametrizations.weight.original1"]


def test_param_grads_at_bench_config(golden):
    """128+128/4 (BASELINE configs[3], the bench workload): parameter gradients of the render loss at the
    reference's sample positions vs the reference's own autograd."""
    g = golden("render_pert_128_128_4")
    net, var, beta, r = build(10, True, n_samples=128, n_importance=128, up_sample_steps=4)
    B = g["rays_o"].shape[0]
    lin = torch.linspace(-1, 1, B, device=dev).reshape(B, 1)
    o = _render_core_at_reference_z(g, net, var, beta, r)
    depth = o["depth"] * g["depth_scale"].to(dev)
    loss = (torch.nn.functional.mse_loss(o["edge"], g["true_edge"].to(dev)) + 0.01 * o["gradient_error_near_surface"]
            + 0.1 * o["gradient_error"] + 0.05 * (depth * lin).mean() + 0.05 * (o["normals"] * g["rays_o"].to(dev)).mean())
    assert abs(float(loss) - float(g["loss"])) <= 5e-4 * max(1.0, abs(float(g["loss"])))
    loss.backward()
    for n, p in net.named_parameters():
        ref = g[f"dloss.{n}"]
        assert maxdiff(p.grad.cpu(), ref) <= 2e-3 * (float(ref.abs().max()) + 1e-12), n      # measured 8.3e-4
        assert l2rel(p.grad, ref) <= 1.5e-3, n                                               # measured 6.0e-4
    for name, p in (("variance", var.variance), ("beta", beta.beta), ("gamma", beta.gamma)):
        ref = g[f"dloss.{name}"]
        assert maxdiff(p.grad.cpu(), ref) <= 2e-3 * (float(ref.abs().max()) + 1e-9) + 1e-9, name


@pytest.mark.timeout(300)
def test_eikonal_gradient_at_production_scale():
    """B = 1024 rays x 128 samples, loss = 0.01 * gradient_error alone: d loss / d grad is ~1e-8 per component,
    below fp16's subnormal range -- the device-side loss scaling must carry it through the fp16 backward.
    Reference: the fp64 oracle (autograd through the restated reference path)."""
    from oracle import emap_oracle as O
    from tests.conftest import load_golden
    B, n = 1024, 128
    net, var, beta, r = build(10, True, n_samples=n, n_importance=0, up_sample_steps=4)
    o_c, d_c = O.synthetic_rays(B)
    near, far, ds = torch.full((B, 1), 0.05), torch.full((B, 1), 6.0), torch.ones(B, 1)
    t_rand = O.synthetic_t_rand(B)
    p = O.perturbed_params(O.UDFParams.from_state_dict(load_golden("net_init_state"))).to(torch.float64)
    p.requires_grad_(True)
    s = O.ScalarParams(*(torch.tensor([v], dtype=torch.float64) for v in (0.3, 0.5, 0.3)))
    cfg = O.RenderConfig(n_samples=n, n_importance=0, up_sample_steps=4)
    ro = O.render(p, s, cfg, o_c.double(), d_c.double(), near.double(), far.double(), ds.double(),
                  cos_anneal_ratio=1.0, flip_saturation=0.9, t_rand=t_rand.double())
    ref = torch.autograd.grad(0.01 * ro["gradient_error"], p.tensors())
    torch.manual_seed(7)
    out = r.render(o_c.to(dev), d_c.to(dev), near.to(dev), far.to(dev), ds.to(dev), cos_anneal_ratio=1.0,
                   flip_saturation=0.9)
    loss = 0.01 * out["gradient_error"]
    assert abs(float(loss) - 0.01 * float(ro["gradient_error"])) <= 1e-3 * abs(0.01 * float(ro["gradient_error"]))
    loss.backward()
    r.check_numerics()
    for (name, prm), gr in zip(net.named_parameters(), ref):
        # measured on B200: 7.1e-4 worst (max norm 9.0e-4); the unscaled backward of round 1: 1.0 (all lost)
        assert l2rel(prm.grad, gr) <= 2e-3, (name, l2rel(prm.grad, gr))
        assert maxdiff(prm.grad.cpu(), gr) <= 2e-3 * float(gr.abs().max()), name


def test_bf16_network_forward_and_gradients(golden):
    """BASELINE config C2 trains in bf16: single bf16 MMA forward (8-bit mantissa), fp16 loss-scaled backward."""
    g = golden("mlp_pert")
    net, var, beta, r = build(10, True, precision="bf16", n_samples=64, n_importance=0, up_sample_steps=5)
    x = g["x"].to(dev)
    y, _ = net(x)
    gg = net.gradient(x.clone()).squeeze(1)
    assert maxdiff(y.detach().cpu(), g["out"]) <= 2e-2                   # measured 8.7e-3 (bf16: 8-bit mantissa)
    # d udf/dx: a CPU emulation of single-bf16-MMA arithmetic (weights and layer inputs rounded to bf16) gives
    # 2.5e-3 median / 9e-3 max against fp32 on these points
    dg = (gg.detach().cpu() - g["grad"].reshape(-1, 3)).abs().max(dim=1)[0]
    print("bf16 grad error: median %.3e  p90 %.3e  max %.3e" % tuple(float(torch.quantile(dg, q)) for q in (0.5, 0.9, 1.0)))
    assert float(dg.median()) <= 1e-2
    assert float(dg.max()) <= 6e-2
    loss = (g["cu"].to(dev) * y).sum() + (g["cg"].to(dev) * gg).sum()
    loss.backward()
    for n, p in net.named_parameters():
        ref = g[f"dgrad.{n}"]
        assert l2rel(p.grad, ref) <= 6e-2, (n, l2rel(p.grad, ref))         # measured 2.9e-2
    # render(): BASELINE config C2's sizes, end to end (measured: z 4.1e-4, edge 7.4e-4, depth 1.3e-3)
    gr = golden("render_pert_64_64_4")
    netb, varb, betab, rb = build(10, True, precision="bf16", n_samples=64, n_importance=64, up_sample_steps=4)
    with torch.no_grad():
        ob = run_render(gr, rb)
    assert maxdiff(ob["edge"].cpu(), gr["out.edge"]) <= 2e-3
    assert maxdiff(ob["weight_sum"].cpu(), gr["out.weight_sum"]) <= 2e-3
    assert maxdiff(ob["mid_z_vals"].cpu(), gr["out.mid_z_vals"]) <= 2e-3
    assert maxdiff(ob["depth"].cpu(), gr["out.depth"]) <= 5e-3


def test_status_word_replaces_pdb_nan_guards():
    """NaN among the new samples of an up-sampling step (reference: pdb.set_trace() at
    udf_renderer_blending.py:102-107 / :346-351) -> device flag -> FloatingPointError; then cleared."""
    from emap_b200 import ops
    ops.check_status(dev)
    bins = torch.sort(torch.rand(8, 16, device=dev), -1)[0]
    w = torch.rand(8, 15, device=dev)
    ops.sample_pdf_det(bins, w, 8)
    ops.check_status(dev)                                   # clean input: no flag
    w[3, 4] = float("nan")
    ops.sample_pdf_det(bins, w, 8)
    with pytest.raises(FloatingPointError, match="up-sampling"):
        ops.check_status(dev)
    ops.check_status(dev)                                   # cleared by the raise
    # the non-blocking poll raises at the latest at the call after the copy has landed
    ops.sample_pdf_det(bins, w, 8)
    ops.poll_status(dev)
    torch.cuda.synchronize()
    with pytest.raises(FloatingPointError):
        ops.poll_status(dev)


def test_validate_loop_and_checkpoint_round_trip(tmp_path):
    """the reference's validate() (runner_udf.py:287-408: a full image rendered in batch_size
    chunks, ragged last chunk, scalar near/far, perturb on, autograd on) and its checkpoint format
    (runner_udf.py:252-285) driven through the drop-in classes; images vs the CPU oracle chunk by chunk."""
    from emap_b200.udf_model import BetaNetwork, SingleVarianceNetwork, UDFNetwork
    from emap_b200.udf_renderer_blending import UDFRendererBlending
    from oracle import emap_oracle as O
    net, var, beta, r = build(10, True, n_samples=64, n_importance=50, up_sample_steps=5)
    H, W, batch_size = 18, 24, 100
    o, d = O.synthetic_rays(H * W, seed=5)
    rays_o, rays_d = o.reshape(H, W, 3).to(dev), d.reshape(H, W, 3).to(dev)
    depth_scale = torch.linspace(0.8, 1.2, H * W).reshape(H, W, 1).to(dev)
    near, far = 0.05, 6.0

    def validate(renderer):
        ro = rays_o.reshape(-1, 3).split(batch_size)
        rd = rays_d.reshape(-1, 3).split(batch_size)
        dsc = depth_scale.reshape(-1, 1).split(batch_size)
        out_edge, out_depth, out_normal = [], [], []
        for ro_b, rd_b, ds_b in zip(ro, rd, dsc):
            render_out = renderer.render(ro_b, rd_b, near, far, depth_scale=ds_b, color_maps=None, pose=None,
                                         fx=1.0, fy=1.0, cos_anneal_ratio=1.0, background_rgb=None)
            out_edge.append(render_out["edge"].detach().cpu())
            out_depth.append(render_out["depth"].detach().cpu())
            nsmp = renderer.n_samples + renderer.n_importance
            out_normal.append((render_out["gradients_flip"] * render_out["weights"][:, :nsmp, None]).sum(dim=1)
                              .detach().cpu())
            del render_out
        return (torch.cat(out_edge).reshape(H, W), torch.cat(out_depth).reshape(H, W),
                torch.cat(out_normal).reshape(H, W, 3))

    torch.manual_seed(123)
    edge, depth, normal = validate(r)
    r.check_numerics()
    # the oracle, chunk by chunk, with the draws the global CPU generator handed to render()
    p = O.UDFParams.from_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
    s = O.ScalarParams(var.variance.detach().cpu(), beta.beta.detach().cpu(), beta.gamma.detach().cpu())
    cfg = O.RenderConfig(n_samples=64, n_importance=50, up_sample_steps=5)
    torch.manual_seed(123)
    e_ref, d_ref = [], []
    for ro_b, rd_b, ds_b in zip(o.split(batch_size), d.split(batch_size), depth_scale.reshape(-1, 1).cpu().split(batch_size)):
        Bc = ro_b.shape[0]
        t_rand = torch.rand(Bc, 1) - 0.5
        with torch.no_grad():
            ref = O.render(p, s, cfg, ro_b, rd_b, torch.full((Bc, 1), near), torch.full((Bc, 1), far), ds_b,
                           cos_anneal_ratio=1.0, flip_saturation=0.0, t_rand=t_rand)
        e_ref.append(ref["edge"]); d_ref.append(ref["depth"])
    assert maxdiff(edge.reshape(-1, 1), torch.cat(e_ref)) <= 2e-3
    assert maxdiff(depth.reshape(-1, 1), torch.cat(d_ref)) <= 1.2e-2
    assert torch.isfinite(normal).all()

    # ---- checkpoint: the reference's dict, its key names, torch.save / torch.load, optimizer state included
    params = list(net.parameters()) + list(var.parameters()) + list(beta.parameters())
    opt = torch.optim.Adam(params, lr=1e-4)
    torch.manual_seed(5)
    out = r.render(rays_o.reshape(-1, 3)[:64], rays_d.reshape(-1, 3)[:64], near, far, depth_scale.reshape(-1, 1)[:64],
                   cos_anneal_ratio=1.0)
    (out["edge"].mean() + 0.1 * out["gradient_error"]).backward()
    opt.step()
    ckpt = {"udf_network_fine": net.state_dict(), "variance_network_fine": var.state_dict(),
            "beta_network": beta.state_dict(), "optimizer": opt.state_dict(), "iter_step": 1}
    path = os.path.join(tmp_path, "ckpt_best.pth")
    torch.save(ckpt, path)
    expect = set(NAMES)
    assert set(ckpt["udf_network_fine"].keys()) == expect                       # SURVEY §5 key names
    assert set(ckpt["variance_network_fine"].keys()) == {"variance", "second_variance"}
    assert set(ckpt["beta_network"].keys()) == {"beta", "gamma", "zeta"}
    torch.manual_seed(9)
    before, _, _ = validate(r)

    torch.manual_seed(1)                        # a differently initialised set of modules, then load
    net2 = UDFNetwork(d_in=3, d_out=1, d_hidden=256, n_layers=8, skip_in=[4], multires=10, bias=0.5, scale=1.0,
                      geometric_init=True, weight_norm=True, udf_type="abs").to(dev)
    var2, beta2 = SingleVarianceNetwork(0.1).to(dev), BetaNetwork(0.1, 0.1, 0.1, 5e-5, True, True, False).to(dev)
    opt2 = torch.optim.Adam(list(net2.parameters()) + list(var2.parameters()) + list(beta2.parameters()), lr=1e-4)
    ck = torch.load(path, map_location=dev)
    net2.load_state_dict(ck["udf_network_fine"])
    var2.load_state_dict(ck["variance_network_fine"])
    beta2.load_state_dict(ck["beta_network"])
    opt2.load_state_dict(ck["optimizer"])
    assert ck["iter_step"] == 1
    r2 = UDFRendererBlending(None, net2, var2, beta2, n_samples=64, n_importance=50, n_outside=0,
                             up_sample_steps=5, perturb=1.0, device=dev)
    torch.manual_seed(9)
    after, _, _ = validate(r2)
    assert torch.equal(before, after)                                           # bit-identical re-render
