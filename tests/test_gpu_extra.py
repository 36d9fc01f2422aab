"""GPU: next-row coverage and edge cases through the public classes / C ABI.

* grid UDF query + normalised gradient as `extract_edge` issues them
  (reference: src/edge_extraction/extract_pointcloud.py:5-95, runner_udf.py:520-526), 4096-point batches;
* udf_type in {square, sdf} and scale != 1 (udf_model.py:82-88, :91, :108), forward + backward;
* ragged / extreme sizes: 1 point, 1 ray, ray counts not a multiple of the block, 512 samples per ray
  (the kernels' maximum), limits rejected loudly;
* perturb_overwrite=0 with tensor near/far, background_rgb, bf16 operands end to end.
"""
import pytest
import torch

from oracle import emap_oracle as O
from tests.helpers import maxdiff, oracle_params
from tests.test_gpu_render import build

pytestmark = pytest.mark.gpu
dev = "cuda"


def test_grid_query_like_extract_edge():
    net, var, beta, r = build(10, True, n_samples=64, n_importance=0, up_sample_steps=5)
    p = oracle_params(True)
    N = 24
    lin = torch.linspace(-1.0, 1.0, N)
    grid = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)
    func = lambda pts: net.udf(pts)[0]                                     # runner_udf.py:520
    def func_grad(pts):                                                    # runner_udf.py:522-526
        g = net.gradient(pts).squeeze(1)
        return g / (torch.linalg.norm(g, dim=-1, keepdim=True) + 1e-12)
    udf, grad = [], []
    with torch.no_grad():
        for chunk in torch.split(grid, 4096):                              # max_batch=4096
            udf.append(func(chunk.to(dev)).cpu())
    for chunk in torch.split(grid, 4096):
        grad.append(func_grad(chunk.to(dev).clone()).detach().cpu())
    udf, grad = torch.cat(udf), torch.cat(grad)
    ref_u = O.udf_forward(p, grid)[0][:, :1]
    ref_g = O.udf_gradient(p, grid).detach()
    ref_g = ref_g / (ref_g.norm(dim=-1, keepdim=True) + 1e-12)
    assert udf.shape == (N ** 3, 1)
    assert maxdiff(udf, ref_u) <= 5e-5 * max(1.0, float(ref_u.max()))
    assert maxdiff(grad, ref_g) <= 2e-4


@pytest.mark.parametrize("udf_type,scale", [("square", 1.0), ("sdf", 1.0), ("abs", 2.0)])
def test_udf_type_and_scale(udf_type, scale):
    from emap_b200.udf_model import UDFNetwork
    torch.manual_seed(0)
    net = UDFNetwork(d_in=3, d_out=1, d_hidden=256, n_layers=8, skip_in=[4], multires=10, bias=0.5,
                     scale=scale, udf_type=udf_type)
    p = O.perturbed_params(O.UDFParams.from_state_dict(net.state_dict()))
    sd = net.state_dict()
    for l in range(9):
        sd[f"lin{l}.parametrizations.weight.original1"] = p.v[l]
        sd[f"lin{l}.parametrizations.weight.original0"] = p.g[l]
        sd[f"lin{l}.bias"] = p.b[l]
    net.load_state_dict(sd)
    net = net.to(dev)
    p.scale, p.udf_type = scale, udf_type
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(300, 3, generator=g) * 2 - 1) * (1.2 / scale)
    cu, cg = torch.randn(300, 1, generator=g), torch.randn(300, 3, generator=g)
    y, _ = net(x.to(dev))
    gg = net.gradient(x.to(dev).clone()).squeeze(1)
    p.requires_grad_(True)
    yr = O.udf_forward(p, x)[0]
    gr = O.udf_gradient(p, x)
    assert maxdiff(y.cpu(), yr) <= 1e-4 * max(1.0, float(yr.abs().max()))
    assert maxdiff(gg.cpu(), gr) <= 1e-4 * max(1.0, float(gr.abs().max()))
    loss = (cu.to(dev) * y).sum() + (cg.to(dev) * gg).sum()
    net.zero_grad()
    loss.backward()
    lr = (cu * yr).sum() + (cg * gr).sum()
    ref = torch.autograd.grad(lr, p.tensors())
    for (n, q), rg in zip(net.named_parameters(), ref):
        assert maxdiff(q.grad.cpu(), rg) <= 1.5e-2 * (float(rg.abs().max()) + 1e-9), n


def test_ragged_and_extreme_sizes():
    from emap_b200 import ops, _cabi as C
    net, var, beta, r = build(10, True, n_samples=448, n_importance=64, up_sample_steps=1)
    p = oracle_params(True)
    one = torch.tensor([[0.1, -0.2, 0.3]])
    u, _ = net(one.to(dev))
    assert maxdiff(u.cpu(), O.udf_forward(p, one)[0]) <= 5e-5
    # 512 samples per ray (kernel maximum), 5 rays (not a multiple of the 4-ray block)
    B = 5
    o, d = O.synthetic_rays(B)
    near, far = torch.full((B, 1), 0.05), torch.full((B, 1), 6.0)
    torch.manual_seed(7)
    out = r.render(o.to(dev), d.to(dev), near.to(dev), far.to(dev), torch.ones(B, 1, device=dev),
                   cos_anneal_ratio=1.0)
    assert out["weights"].shape == (B, 512) and torch.isfinite(out["weights"]).all()
    mz = out["mid_z_vals"]
    assert bool((mz[:, 1:] >= mz[:, :-1]).all())
    # one ray
    out1 = r.render(o[:1].to(dev), d[:1].to(dev), near[:1].to(dev), far[:1].to(dev), torch.ones(1, 1, device=dev))
    assert out1["edge"].shape == (1, 1)
    # limits are rejected loudly, not truncated
    r2 = build(10, True, n_samples=512, n_importance=64, up_sample_steps=1)[3]
    with pytest.raises(RuntimeError, match="512 samples"):
        r2.render(o.to(dev), d.to(dev), near.to(dev), far.to(dev), torch.ones(B, 1, device=dev))
    r3 = build(10, True, n_samples=64, n_importance=65, up_sample_steps=1)[3]
    with pytest.raises(RuntimeError, match="64 new samples"):
        r3.render(o.to(dev), d.to(dev), near.to(dev), far.to(dev), torch.ones(B, 1, device=dev))


def test_perturb_overwrite_background_and_bf16(golden):
    g = golden("render_pert_64_0")
    net, var, beta, r = build(10, True, n_samples=64, n_importance=0, up_sample_steps=5)
    B = g["rays_o"].shape[0]
    args = (g["rays_o"].to(dev), g["rays_d"].to(dev), g["near"].to(dev), g["far"].to(dev),
            g["depth_scale"].to(dev))
    a = r.render(*args, cos_anneal_ratio=None, perturb_overwrite=0, flip_saturation=0.9)
    p = oracle_params(True)
    from tests.helpers import oracle_scalars
    cfg = O.RenderConfig(n_samples=64, n_importance=0, perturb=0.0)
    ref = O.render(p, oracle_scalars(), cfg, g["rays_o"], g["rays_d"], g["near"], g["far"], g["depth_scale"],
                   cos_anneal_ratio=None, flip_saturation=0.9, t_rand=None)
    assert maxdiff(a["mid_z_vals"].cpu(), ref["mid_z_vals"]) == 0.0
    assert maxdiff(a["weights"].cpu(), ref["weights"]) <= 2e-4
    with pytest.raises(ValueError):
        r.render(args[0], args[1], 0.05, 6.0, args[4], perturb_overwrite=0)   # reference crashes here too
    bg = torch.tensor([[0.25]], device=dev)
    b = r.render(*args, cos_anneal_ratio=None, perturb_overwrite=0, flip_saturation=0.9, background_rgb=bg)
    assert maxdiff(b["edge"], a["edge"] + 0.25 * (1 - a["weights"].sum(-1, keepdim=True))) <= 1e-6
    # bf16 operands end to end: finite, close (8-bit mantissa)
    nb = build(10, True, precision="bf16", n_samples=64, n_importance=0, up_sample_steps=5)[3]
    c = nb.render(*args, cos_anneal_ratio=None, perturb_overwrite=0, flip_saturation=0.9)
    assert torch.isfinite(c["weights"]).all()
    assert maxdiff(c["edge"].cpu(), ref["edge"]) <= 5e-3        # flat sampling; hierarchical: tests/test_gpu_parity_r2.py


def test_rendering_network_standalone_operator(golden):
    """RenderingNetwork.forward (dead code in the reference, built as a standalone operator):
    same seed -> the reference's own initial weights and state_dict keys; output vs the reference fixture."""
    from emap_b200.udf_model import RenderingNetwork
    g = golden("rendering_network")
    torch.manual_seed(3)
    rn = RenderingNetwork(d_feature=256, mode="no_normal", d_in=6, d_out=1, d_hidden=128, n_layers=4,
                          weight_norm=True, multires_view=4, squeeze_out=True)
    sd = rn.state_dict()
    ref_sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    assert set(sd.keys()) == set(ref_sd.keys())
    for k in sd:
        assert torch.equal(sd[k], ref_sd[k]), k                    # bit-identical initialisation
    rn = rn.to(dev)
    c = rn(g["pts"].to(dev), g["normals"].to(dev), g["view_dirs"].to(dev), g["feat"].to(dev))
    assert c.shape == g["color"].shape and c.dtype == torch.float32
    assert maxdiff(c.cpu(), g["color"]) <= 2e-6                    # fp32 sums in a different order than the CPU GEMM
    # the other two input modes against the oracle restatement
    for mode, d_in in (("idr", 12), ("no_view_dir", 9)):       # d_in counts points, view, normals, -normals
        torch.manual_seed(4)
        r2 = RenderingNetwork(d_feature=256, mode=mode, d_in=d_in, d_out=3, d_hidden=64, n_layers=2,
                              weight_norm=True, multires_view=4 if mode == "idr" else 0, squeeze_out=False)
        W = [getattr(r2, f"lin{l}").weight.detach() for l in range(3)]
        b = [getattr(r2, f"lin{l}").bias.detach() for l in range(3)]
        ref = O.rendering_network_forward(W, b, mode, g["pts"], g["normals"], g["view_dirs"], g["feat"],
                                          multires_view=4 if mode == "idr" else 0, squeeze_out=False, d_out=3)
        r2 = r2.to(dev)
        got = r2(g["pts"].to(dev), g["normals"].to(dev), g["view_dirs"].to(dev), g["feat"].to(dev))
        assert maxdiff(got.cpu(), ref) <= 1e-5 * max(1.0, float(ref.abs().max())), mode
