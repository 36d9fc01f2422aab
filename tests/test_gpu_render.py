"""GPU parity of the drop-in classes (UDFNetwork / UDFRendererBlending.render) against the fixtures
generated from the reference, end to end through the C ABI.

End-to-end tolerances are looser than the per-stage ones because the hierarchical sampler amplifies
MLP rounding: the last up-sampling step evaluates sigmoid(1024*udf) and a 2048-sharp logistic, so
an fp32-class udf error of 2e-5 moves a few new samples in z.  Bounds here (fp32x3 mode, vs the fp32 CPU
reference) = 2x the values measured on B200 (profiles/r02_parity.json): per-ray edge / weight_sum 2e-4
absolute, depth 1e-3, z 1.5e-3 (2e-2 on the pristine sphere initialisation, see
tests/test_gpu_parity_r2.py, which also compares every per-sample tensor at the reference's positions).
"""
import pytest
import torch

from tests.conftest import load_golden
from tests.helpers import maxdiff

pytestmark = pytest.mark.gpu
dev = "cuda"


def build(multires=10, pert=True, precision="fp32", **rkw):
    from emap_b200.udf_model import BetaNetwork, SingleVarianceNetwork, UDFNetwork
    from emap_b200.udf_renderer_blending import UDFRendererBlending
    from oracle import emap_oracle as O
    torch.manual_seed(0)
    net = UDFNetwork(d_in=3, d_out=1, d_hidden=256, n_layers=8, skip_in=[4], multires=multires, bias=0.5,
                     scale=1.0, geometric_init=True, weight_norm=True, udf_type="abs", precision=precision)
    sd = load_golden("net_init_state" if multires == 10 else "net_init_state_mr6")
    init_sd = {k: v.clone() for k, v in net.state_dict().items()}
    for k in sd:                       # same seed -> bit-identical init as the reference constructor
        assert torch.equal(init_sd[k], sd[k]), k
    if pert:
        p2 = O.perturbed_params(O.UDFParams.from_state_dict(sd, multires=multires))
        for l in range(9):
            sd[f"lin{l}.parametrizations.weight.original1"] = p2.v[l]
            sd[f"lin{l}.parametrizations.weight.original0"] = p2.g[l]
            sd[f"lin{l}.bias"] = p2.b[l]
    net.load_state_dict(sd)
    net = net.to(dev)
    var = SingleVarianceNetwork(0.3).to(dev)
    beta = BetaNetwork(0.5, 0.3, 0.3, 5e-5, True, True, False).to(dev)
    r = UDFRendererBlending(None, net, var, beta, n_outside=0, perturb=1.0, device=dev, **rkw)
    return net, var, beta, r


CASES = [
    ("init_64_50_5", False, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5)),
    ("pert_64_64_4", True, 10, dict(n_samples=64, n_importance=64, up_sample_steps=4)),
    ("pert_64_0", True, 10, dict(n_samples=64, n_importance=0, up_sample_steps=5)),
    ("pert_128_128_4", True, 10, dict(n_samples=128, n_importance=128, up_sample_steps=4)),
    ("mr6_64_50_5", True, 6, dict(n_samples=64, n_importance=50, up_sample_steps=5)),
    ("var_biased", True, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5, use_unbias_render=False)),
    ("var_theorical", True, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5, sdf2alpha_type="theorical")),
    ("var_normgrad", True, 10, dict(n_samples=64, n_importance=50, up_sample_steps=5, use_norm_grad_for_cosine=True)),
    ("var_mix", True, 10, dict(n_samples=64, n_importance=60, up_sample_steps=5, upsampling_type="mix")),
]


def run_render(g, r):
    car = float(g["cos_anneal_ratio"])
    torch.manual_seed(7)      # render() draws rand([B,1]) from the global CPU generator, like the reference
    return r.render(g["rays_o"].to(dev), g["rays_d"].to(dev), g["near"].to(dev), g["far"].to(dev),
                    g["depth_scale"].to(dev), cos_anneal_ratio=None if car < 0 else car,
                    flip_saturation=float(g["flip_saturation"]))


@pytest.mark.parametrize("tag,pert,multires,rkw", CASES)
def test_render_matches_reference(golden, tag, pert, multires, rkw):
    g = golden(f"render_{tag}")
    net, var, beta, r = build(multires, pert, **rkw)
    out = run_render(g, r)
    torch.cuda.synchronize()
    B, n = g["out.udf"].shape
    keys = ["udf", "edge", "weight_sum", "weight_sum_fg_bg", "depth", "variance", "beta", "gamma", "normals",
            "gradients", "gradients_flip", "weights", "gradient_error", "gradient_error_near_surface",
            "inside_sphere", "gradient_mag", "mid_z_vals", "dists"]
    assert sorted(out.keys()) == sorted(keys)
    for k in keys:
        if k == "variance":
            assert out[k].shape == (B * n, 1)
            assert maxdiff(out[k][:1].cpu(), g["out.variance0"]) <= 1e-6 * float(g["out.variance0"])
            continue
        ref = g[f"out.{k}"]
        assert out[k].shape == ref.shape, (k, out[k].shape, ref.shape)
        assert out[k].dtype == torch.float32 and out[k].is_cuda
    o = {k: v.detach().cpu() for k, v in out.items()}
    flat = rkw["n_importance"] == 0
    ztol = 2e-6 if flat else (1.5e-3 if pert else 2e-2)
    assert maxdiff(o["mid_z_vals"], g["out.mid_z_vals"]) <= ztol, maxdiff(o["mid_z_vals"], g["out.mid_z_vals"])
    assert maxdiff(o["edge"], g["out.edge"]) <= 2e-4
    assert maxdiff(o["weight_sum"], g["out.weight_sum"]) <= 2e-4
    assert maxdiff(o["depth"], g["out.depth"]) <= 1e-3
    assert abs(float(o["gradient_error"]) - float(g["out.gradient_error"])) <= 2e-3 * max(1, float(g["out.gradient_error"]))
    assert float(o["beta"]) == pytest.approx(float(g["out.beta"]), rel=1e-6)
    assert float(o["gamma"]) == pytest.approx(float(g["out.gamma"]), rel=1e-6)
    if flat:   # identical sample positions -> per-sample tensors comparable one to one
        assert maxdiff(o["udf"], g["out.udf"]) <= 5e-5 * 3
        assert maxdiff(o["weights"], g["out.weights"]) <= 2e-4
        assert maxdiff(o["gradients"], g["out.gradients"]) <= 1e-4
        assert maxdiff(o["normals"], g["out.normals"]) <= 5e-4


def test_importance_sampling_positions(golden):
    from emap_b200 import ops
    g = golden("upsample_pert_64_64_4")
    net, var, beta, r = build(10, True, n_samples=64, n_importance=64, up_sample_steps=4)
    sd = torch.tensor([float(g["sample_dist"])], device=dev)
    z = r.importance_sample(g["rays_o"].to(dev), g["rays_d"].to(dev), g["z0"].to(dev), sd)
    assert z.shape == g["z_final"].shape
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    assert maxdiff(z.cpu(), g["z_final"]) <= 1.5e-3


def test_properties_at_scale():
    """size-independent properties at the bench size (4096 rays x 128+128): weights in [0,1],
    sum <= 1, z sorted and inside [z0, z_last], finite outputs, ray-permutation equivariance."""
    from oracle import emap_oracle as O
    net, var, beta, r = build(10, True, n_samples=128, n_importance=128, up_sample_steps=4)
    B = 4096
    o, d = O.synthetic_rays(B)
    near, far = torch.full((B, 1), 0.05), torch.full((B, 1), 6.0)
    ds = torch.ones(B, 1)
    torch.manual_seed(7)
    with torch.no_grad():
        out = r.render(o.to(dev), d.to(dev), near.to(dev), far.to(dev), ds.to(dev), cos_anneal_ratio=1.0,
                       flip_saturation=0.9)
        perm = torch.randperm(B)
        torch.manual_seed(7)
        t = torch.rand(B, 1)
        # same jitter per ray after permutation: feed it by seeding and un-permuting is not possible
        # through the public API, so check equivariance on the flat (perturb-free) path instead
        r2 = build(10, True, n_samples=128, n_importance=128, up_sample_steps=4)[3]
        r2.perturb = 0
        a = r2.render(o.to(dev), d.to(dev), near.to(dev), far.to(dev), ds.to(dev), cos_anneal_ratio=1.0)
        b = r2.render(o[perm].to(dev), d[perm].to(dev), near.to(dev), far.to(dev), ds.to(dev), cos_anneal_ratio=1.0)
    w = out["weights"]
    assert w.shape == (B, 256) and torch.isfinite(w).all()
    assert float(w.min()) >= 0.0 and float(w.max()) <= 1.0 + 1e-6
    assert float(out["weight_sum"].max()) <= 1.0 + 1e-4
    mz = out["mid_z_vals"]
    assert bool((mz[:, 1:] >= mz[:, :-1]).all())
    for k, v in out.items():
        assert torch.isfinite(v).all(), k
    assert torch.equal(a["weights"][perm.to(dev)], b["weights"])
    assert torch.equal(a["edge"][perm.to(dev)], b["edge"])


def test_properties_c5_sweep_size():
    """BASELINE config C5 (65536 rays x 256 flat samples, 16.8 M points per launch): size-independent
    properties -- finite outputs, weights in [0,1] summing to <= 1, sorted samples, and shard invariance:
    rendering the batch in two halves (what the multi-GPU path does with the rays) is bit-identical to
    rendering it whole."""
    from oracle import emap_oracle as O
    net, var, beta, r = build(10, True, n_samples=256, n_importance=0, up_sample_steps=4)
    r.perturb = 0
    B = 65536
    o, d = O.synthetic_rays(B)
    near, far = torch.full((B, 1), 0.05).to(dev), torch.full((B, 1), 6.0).to(dev)
    ds = torch.ones(B, 1).to(dev)
    o, d = o.to(dev), d.to(dev)
    with torch.no_grad():
        full = r.render(o, d, near, far, ds, cos_anneal_ratio=1.0, flip_saturation=0.9)
        h = B // 2
        lo = r.render(o[:h], d[:h], near[:h], far[:h], ds[:h], cos_anneal_ratio=1.0, flip_saturation=0.9)
        hi = r.render(o[h:], d[h:], near[h:], far[h:], ds[h:], cos_anneal_ratio=1.0, flip_saturation=0.9)
    w = full["weights"]
    assert w.shape == (B, 256)
    for k in ("weights", "edge", "depth", "normals", "udf", "gradients", "gradient_mag"):
        assert torch.isfinite(full[k]).all(), k
    assert float(w.min()) >= 0.0 and float(w.max()) <= 1.0 + 1e-6
    assert float(full["weight_sum"].max()) <= 1.0 + 1e-4
    mz = full["mid_z_vals"]
    assert bool((mz[:, 1:] >= mz[:, :-1]).all())
    for k in ("weights", "edge", "depth", "normals", "udf", "gradients"):
        assert torch.equal(full[k], torch.cat([lo[k], hi[k]])), k
