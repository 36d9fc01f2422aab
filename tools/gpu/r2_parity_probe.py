"""Round-2 parity probe: measures every error the GPU parity tests bound, so that the bounds in tests/ are the
measured values with a stated margin ("assert what you measure").  Writes gpurun_out/r2_parity.json.
usage: python tools/gpu/r2_parity_probe.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops  # noqa: E402
from oracle import emap_oracle as O  # noqa: E402
from tests.conftest import load_golden  # noqa: E402
from tests.helpers import maxdiff  # noqa: E402
from tests.test_gpu_render import CASES, build, run_render  # noqa: E402

dev = "cuda"
out = {}


def l2rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))


def maxrel(a, b):
    return maxdiff(a.cpu(), b.cpu()) / (float(b.double().abs().max()) + 1e-300)


# ---------------------------------------------------------------- 1. render(): end to end and at the reference's z
res = {}
for tag, pert, multires, rkw in CASES:
    g = load_golden(f"render_{tag}")
    net, var, beta, r = build(multires, pert, **rkw)
    with torch.no_grad():
        o = run_render(g, r)
    e = {k: maxdiff(o[k].cpu(), g[f"out.{k}"]) for k in ("mid_z_vals", "edge", "weight_sum", "depth", "normals")}
    e["gradient_error_rel"] = abs(float(o["gradient_error"]) - float(g["out.gradient_error"])) / max(1.0, float(g["out.gradient_error"]))
    z_ref = (g["out.mid_z_vals"] - 0.5 * g["out.dists"]).to(dev)
    sd = torch.tensor([float(g["out.dists"][0, -1])], device=dev)
    car = float(g["cos_anneal_ratio"])
    with torch.no_grad():
        c = r.render_core(g["rays_o"].to(dev), g["rays_d"].to(dev), z_ref, sd, net, var, beta_network=beta,
                          cos_anneal_ratio=None if car < 0 else car, flip_saturation=float(g["flip_saturation"]))
    f = {k: maxdiff(c[k].cpu(), g[f"out.{k}"]) for k in ("udf", "weights", "gradients", "gradients_flip",
                                                         "gradient_mag", "edge", "normals", "mid_z_vals", "dists")}
    f["depth"] = maxdiff((c["depth"] * g["depth_scale"].to(dev)).cpu(), g["out.depth"])
    f["inside_sphere_equal"] = bool(torch.equal(c["inside_sphere"].cpu(), g["out.inside_sphere"]))
    f["gradient_error_rel"] = abs(float(c["gradient_error"]) - float(g["out.gradient_error"])) / max(1.0, float(g["out.gradient_error"]))
    res[tag] = {"end_to_end": e, "at_reference_z": f}
    print(tag, json.dumps(res[tag]), flush=True)
out["render"] = res

# ---------------------------------------------------------------- 2. up-sampling steps in isolation: flips are knife edges
res = {}
for tag, n0, ni, steps in (("init_64_50_5", 64, 50, 5), ("pert_64_64_4", 64, 64, 4), ("pert_128_128_4", 128, 128, 4)):
    g = load_golden(f"upsample_{tag}")
    o, d = g["rays_o"].to(dev), g["rays_d"].to(dev)
    sd = torch.tensor([float(g["sample_dist"])], device=dev)
    k = ni // steps
    u = torch.linspace(0.5 / k, 1 - 0.5 / k, steps=k)
    per = []
    for i in range(steps):
        zi = g["z0"] if i == 0 else g[f"z{i}"]
        ui = g["udf0"] if i == 0 else g[f"udf{i}"]
        inv_s, bet, gam = O.upsample_schedule(i, steps)
        _, _, z_new, inds, w = ops.upsample_step(o, d, zi.to(dev), ui.to(dev), None, None, u.to(dev), k, sd,
                                                 inv_s, bet, gam, want_inds=True, want_weights=True)
        zr, ir, wr = O.up_sample_unbias(g["rays_o"], g["rays_d"], zi, ui, float(g["sample_dist"]), k,
                                        inv_s, bet, gam, return_aux=True)
        ww = wr.double() + 1e-5
        cdf = torch.cat([torch.zeros(ww.shape[0], 1, dtype=torch.float64), torch.cumsum(ww / ww.sum(-1, keepdim=True), -1)], -1)
        flip = (inds.cpu() != ir).nonzero()
        edges = []
        for ray, j in flip.tolist():
            a, b = int(inds[ray, j]), int(ir[ray, j])
            lo, hi = min(a, b), max(a, b)
            # searchsorted(right=True) differs by one bin <=> a cdf entry in (lo .. hi] sits at the quantile
            edges.append(float((cdf[ray, lo:hi] - float(u[j])).abs().min()))
        dz = (z_new.cpu() - torch.sort(g[f"z_new{i}"], -1)[0]).abs()
        per.append({"w_rel": maxdiff(w.cpu(), wr) / max(1e-3, float(wr.abs().max())), "flips": len(edges),
                    "flip_cdf_minus_u": edges, "z_new_max": float(dz.max()),
                    "z_new_frac_gt_1e-5": float((dz > 1e-5).double().mean()), "z_new_p999": float(dz.flatten().kthvalue(max(1, int(0.999 * dz.numel())))[0])})
    res[tag] = per
    print(tag, json.dumps(per), flush=True)
out["upsample_isolated"] = res

# ---------------------------------------------------------------- 3. parameter gradients vs the reference fixtures
NAMES = []
for _l in range(9):
    NAMES += [f"lin{_l}.bias", f"lin{_l}.parametrizations.weight.original0", f"lin{_l}.parametrizations.weight.original1"]


def grad_errs(named, g, prefix):
    r = {}
    for n, gr in named:
        ref = g[f"{prefix}.{n}"]
        r[n] = (maxrel(gr, ref), l2rel(gr, ref))
    return {"worst_maxrel": max(v[0] for v in r.values()), "worst_l2rel": max(v[1] for v in r.values()),
            "per_tensor": {k: [round(v[0], 6), round(v[1], 6)] for k, v in r.items()}}


res = {}
for modes in (("reverse", "shared"), ("forward", "dual")):
    ops.set_grad_mode(modes[0]); ops.set_backward_mode(modes[1])
    for tag, pert in (("init", False), ("pert", True)):
        g = load_golden(f"mlp_{tag}")
        net, var, beta, r = build(10, pert, n_samples=64, n_importance=0, up_sample_steps=5)
        x = g["x"].to(dev)
        y, _ = net(x)
        gg = net.gradient(x.clone()).squeeze(1)
        loss = (g["cu"].to(dev) * y).sum() + (g["cg"].to(dev) * gg).sum()
        net.zero_grad()
        loss.backward()
        res[f"mlp_{tag}/{modes[0]}+{modes[1]}"] = grad_errs([(n, p.grad) for n, p in net.named_parameters()], g, "dgrad")
    for tag, pert, rkw in (("init_64_50_5", False, dict(n_samples=64, n_importance=50, up_sample_steps=5)),
                           ("pert_64_64_4", True, dict(n_samples=64, n_importance=64, up_sample_steps=4)),
                           ("pert_64_0", True, dict(n_samples=64, n_importance=0, up_sample_steps=5)),
                           ("pert_128_128_4", True, dict(n_samples=128, n_importance=128, up_sample_steps=4))):
        g = load_golden(f"render_{tag}")
        if "loss" not in g:
            continue
        net, var, beta, r = build(10, pert, **rkw)
        B = g["rays_o"].shape[0]
        lin = torch.linspace(-1, 1, B, device=dev).reshape(B, 1)
        z_ref = (g["out.mid_z_vals"] - 0.5 * g["out.dists"]).to(dev)
        sd = torch.tensor([float(g["out.dists"][0, -1])], device=dev)
        car = float(g["cos_anneal_ratio"])
        o = r.render_core(g["rays_o"].to(dev), g["rays_d"].to(dev), z_ref, sd, net, var, beta_network=beta,
                          cos_anneal_ratio=None if car < 0 else car, flip_saturation=float(g["flip_saturation"]))
        depth = o["depth"] * g["depth_scale"].to(dev)
        loss = (torch.nn.functional.mse_loss(o["edge"], g["true_edge"].to(dev)) + 0.01 * o["gradient_error_near_surface"]
                + 0.1 * o["gradient_error"] + 0.05 * (depth * lin).mean() + 0.05 * (o["normals"] * g["rays_o"].to(dev)).mean())
        for m in (net, var, beta):
            m.zero_grad()
        loss.backward()
        e = grad_errs([(n, p.grad) for n, p in net.named_parameters()], g, "dloss")
        e["loss_rel"] = abs(float(loss) - float(g["loss"])) / max(1.0, abs(float(g["loss"])))
        e["scalars"] = {nm: maxrel(p.grad, g[f"dloss.{nm}"]) for nm, p in (("variance", var.variance), ("beta", beta.beta), ("gamma", beta.gamma))}
        res[f"render_{tag}/{modes[0]}+{modes[1]}"] = e
    print(modes, json.dumps({k: (v["worst_maxrel"], v["worst_l2rel"]) for k, v in res.items() if k.endswith(modes[1])}), flush=True)
ops.set_grad_mode(ops.DEFAULT_GRAD_MODE); ops.set_backward_mode(ops.DEFAULT_BWD_MODE)
out["param_grads_fixtures"] = res

# ---------------------------------------------------------------- 4. production scale: eikonal-only and full loss vs the fp64 oracle
res = {}
B, n = 1024, 128
net, var, beta, r = build(10, True, n_samples=n, n_importance=0, up_sample_steps=4)
o_c, d_c = O.synthetic_rays(B)
near, far, ds = torch.full((B, 1), 0.05), torch.full((B, 1), 6.0), torch.ones(B, 1)
te = torch.rand(B, 1, generator=torch.Generator().manual_seed(21))
t_rand = O.synthetic_t_rand(B)
t0 = time.time()
p = O.perturbed_params(O.UDFParams.from_state_dict(load_golden("net_init_state"))).to(torch.float64)
p.requires_grad_(True)
s = O.ScalarParams(torch.tensor([0.3], dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64),
                   torch.tensor([0.3], dtype=torch.float64))
cfg = O.RenderConfig(n_samples=n, n_importance=0, up_sample_steps=4)
ro = O.render(p, s, cfg, o_c.double(), d_c.double(), near.double(), far.double(), ds.double(), cos_anneal_ratio=1.0,
              flip_saturation=0.9, t_rand=t_rand.double())
losses_ref = {"eikonal": 0.01 * ro["gradient_error"],
              "full": torch.nn.functional.mse_loss(ro["edge"], te.double()) + 0.01 * ro["gradient_error_near_surface"] + 0.01 * ro["gradient_error"]}
ref_g = {k: torch.autograd.grad(v, p.tensors(), retain_graph=True) for k, v in losses_ref.items()}
res["oracle_fp64_seconds"] = time.time() - t0
names = [n_ for n_, _ in net.named_parameters()]
for scaling in (True, False):
    os.environ["EMAP_BWD_NOSCALE"] = "0" if scaling else "1"
    for which in ("eikonal", "full"):
        torch.manual_seed(7)
        oo = r.render(o_c.to(dev), d_c.to(dev), near.to(dev), far.to(dev), ds.to(dev), cos_anneal_ratio=1.0, flip_saturation=0.9)
        loss = 0.01 * oo["gradient_error"] if which == "eikonal" else (
            torch.nn.functional.mse_loss(oo["edge"], te.to(dev)) + 0.01 * oo["gradient_error_near_surface"] + 0.01 * oo["gradient_error"])
        for m in (net, var, beta):
            m.zero_grad()
        loss.backward()
        per = {nm: (maxrel(pp.grad, gr), l2rel(pp.grad, gr)) for (nm, pp), gr in zip(net.named_parameters(), ref_g[which])}
        res[f"{which}/{'scaled' if scaling else 'unscaled'}"] = {
            "loss_rel": abs(float(loss) - float(losses_ref[which])) / abs(float(losses_ref[which])),
            "worst_maxrel": max(v[0] for v in per.values()), "worst_l2rel": max(v[1] for v in per.values()),
            "per_tensor": {k: [round(v[0], 6), round(v[1], 6)] for k, v in per.items()}}
        print(which, scaling, res[f"{which}/{'scaled' if scaling else 'unscaled'}"]["worst_maxrel"],
              res[f"{which}/{'scaled' if scaling else 'unscaled'}"]["worst_l2rel"], flush=True)
os.environ["EMAP_BWD_NOSCALE"] = "0"
try:
    ops.check_status(dev)
    res["status_after"] = 0
except FloatingPointError as e:
    res["status_after"] = str(e)
out["production_scale_1024x128"] = res

# ---------------------------------------------------------------- 5. bf16 network (BASELINE config C2)
res = {}
g = load_golden("mlp_pert")
netb, varb, betab, rb = build(10, True, precision="bf16", n_samples=64, n_importance=0, up_sample_steps=5)
x = g["x"].to(dev)
y, _ = netb(x)
gg = netb.gradient(x.clone()).squeeze(1)
res["udf_abs"] = maxdiff(y.detach().cpu(), g["out"])
res["grad_abs"] = maxdiff(gg.detach().cpu(), g["grad"].reshape(-1, 3))
loss = (g["cu"].to(dev) * y).sum() + (g["cg"].to(dev) * gg).sum()
netb.zero_grad()
loss.backward()
res["param_grads"] = grad_errs([(n_, pp.grad) for n_, pp in netb.named_parameters()], g, "dgrad")
res["param_grads"].pop("per_tensor")
gr = load_golden("render_pert_64_64_4")
netb, varb, betab, rb = build(10, True, precision="bf16", n_samples=64, n_importance=64, up_sample_steps=4)
with torch.no_grad():
    ob = run_render(gr, rb)
res["render_e2e"] = {k: maxdiff(ob[k].cpu(), gr[f"out.{k}"]) for k in ("mid_z_vals", "edge", "weight_sum", "depth")}
out["bf16"] = res
print("bf16", json.dumps(res), flush=True)

# ---------------------------------------------------------------- 6. NaN status word
res = {}
try:
    bins = torch.sort(torch.rand(8, 16, device=dev), -1)[0]
    w = torch.rand(8, 15, device=dev)
    w[3, 4] = float("nan")
    ops.sample_pdf_det(bins, w, 8)
    ops.check_status(dev)
    res["nan_samples"] = "NOT raised"
except FloatingPointError as e:
    res["nan_samples"] = str(e)
try:
    ops.check_status(dev)
    res["cleared"] = True
except FloatingPointError as e:
    res["cleared"] = str(e)
out["status_word"] = res
print("status", json.dumps(res), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_parity.json"), "w"), indent=1)
print("wrote gpurun_out/r2_parity.json")
