"""GPU parity: K1r, forward + REVERSE-mode gradient in one tcgen05 kernel (emap_b200/csrc/mlp_rg.cu),
against the reference-generated fixtures, the CPU oracle and the validated forward-mode kernel K1g.

Status: validated on B200 (round 2); K1r + the shared-forward backward are the product default
(ops._GRAD_MODE = "reverse", ops._BWD_MODE = "shared").  The forward-mode kernel K1g and the dual-forward
backward remain as independent cross-checks.

Tolerances: the ones K1g is held to (tests/test_gpu_mlp.py): fp32x3 5e-5 relative to max(1, |ref|max);
fp16 3e-3 / 1e-2.
"""
import pytest
import torch

from tests.helpers import maxdiff, oracle_params

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.timeout(120),
]


def _net(pert, multires=10, elem="fp16", udf_type="abs", scale=1.0):
    from emap_b200 import ops
    p = oracle_params(pert, multires)
    flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
    net = ops.PackedNet(multires, udf_type=udf_type, scale=scale, elem_type=elem)
    net.fold(flat)
    return net, p


def test_stepwise_accumulators_vs_emulation():
    """bring-up diagnostic, run first: the accumulators of tile 0 after each of the 16 MMA steps against the
    float64 emulation of the same algorithm (tests/test_rg_emulation.py) -- names the first step that
    diverges instead of only reporting a wrong gradient."""
    import numpy as np
    from emap_b200 import ops, _cabi as C
    from tests.test_rg_emulation import _emulate
    net, p = _net(True)
    torch.manual_seed(3)
    x = (torch.rand(128, 3) * 2 - 1) * 0.9
    udf, grad, dbg = ops.debug_rgrad(net, C.PREC_FP32X3, x.cuda())
    torch.cuda.synchronize()
    steps = []
    eu, eg = _emulate(p, x, steps=steps)
    assert len(steps) == 16
    got = dbg.cpu().double().numpy()
    for s_i, want in enumerate(steps):
        scale = max(1.0, float(np.abs(want).max()))
        err = float(np.abs(got[s_i] - want).max())
        assert err <= 2e-4 * scale, f"MMA step {s_i}: max |acc - emulation| = {err:.3e} (scale {scale:.2e})"
    assert float(np.abs(udf.cpu().numpy() - eu).max()) <= 5e-5
    assert float(np.abs(grad.cpu().numpy() - eg).max()) <= 5e-5


@pytest.mark.parametrize("pert", [False, True])
@pytest.mark.parametrize("prec,tol_u,tol_g", [("fp32", 5e-5, 5e-5), ("fp16", 3e-3, 1e-2)])
def test_reverse_mode_vs_reference(golden, pert, prec, tol_u, tol_g):
    from emap_b200 import ops, _cabi as C
    g = golden("mlp_pert" if pert else "mlp_init")
    net, _ = _net(pert)
    x = g["x"].cuda()
    udf, grad = ops.udf_forward_grad(net, C.PRECISIONS[prec], pts=x, mode="reverse")
    torch.cuda.synchronize()
    ref_u, ref_g = g["udf"][:, 0], g["grad"][:, 0]
    assert maxdiff(udf.cpu(), ref_u) <= tol_u * max(1.0, float(ref_u.abs().max()))
    assert maxdiff(grad.cpu(), ref_g) <= tol_g * max(1.0, float(ref_g.abs().max()))


def test_reverse_mode_multires6(golden):
    from emap_b200 import ops, _cabi as C
    g = golden("mlp_mr6_pert")
    net, _ = _net(True, multires=6)
    udf, grad = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=g["x"].cuda(), mode="reverse")
    assert maxdiff(udf.cpu(), g["out"][:, 0]) <= 5e-5 * max(1.0, float(g["out"].abs().max()))
    assert maxdiff(grad.cpu(), g["grad"][:, 0]) <= 5e-5 * max(1.0, float(g["grad"].abs().max()))


def test_reverse_mode_rays_ragged_many_tiles_and_repeatable():
    """points given as rays, P not a multiple of 128, several tiles per CTA (the scratch slice and every
    barrier phase are reused), two runs bit-identical, and agreement with K1g far below the tolerance."""
    from emap_b200 import ops, _cabi as C
    from oracle import emap_oracle as O
    net, p = _net(True)
    B, n = 6011, 11                                   # 66,121 points -> 517 tiles on 148 CTAs
    o, d = O.synthetic_rays(B)
    z = torch.rand(B, n) * 3 + 0.5
    args = dict(rays_o=o.cuda(), rays_d=d.cuda(), z=z.cuda())
    u1, g1 = ops.udf_forward_grad(net, C.PREC_FP32X3, mode="reverse", **args)
    u2, g2 = ops.udf_forward_grad(net, C.PREC_FP32X3, mode="reverse", **args)
    uf, gf = ops.udf_forward_grad(net, C.PREC_FP32X3, mode="forward", **args)
    torch.cuda.synchronize()
    assert torch.equal(u1, u2) and torch.equal(g1, g2)
    assert maxdiff(u1, uf) <= 5e-5 and maxdiff(g1, gf) <= 5e-5
    sel = torch.randperm(B * n)[:4096]
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).reshape(-1, 3)[sel]
    ref_u = O.udf_forward(p, pts)[0][:, 0]
    ref_g = O.udf_gradient(p, pts).detach()
    assert maxdiff(u1.cpu()[sel], ref_u) <= 5e-5 * max(1.0, float(ref_u.abs().max()))
    assert maxdiff(g1.cpu()[sel], ref_g) <= 5e-5 * max(1.0, float(ref_g.abs().max()))


def test_split_tail_and_static_schedule_variants_are_bit_identical(golden):
    """rg_flags bit 0: N-split of each step's last K chunk (same MMA order per accumulator column); bit 3 off:
    static round-robin tiles instead of the dynamic counter (the default) -- which CTA runs a tile must not matter;
    bit 2: the instantiation whose MMA issuer walks the 16 steps in a rolled loop (same instruction stream)."""
    from emap_b200 import ops, _cabi as C
    g = golden("mlp_pert")
    net, _ = _net(True)
    x = g["x"].cuda().repeat(60, 1)           # 23,040 points -> 180 tiles: two tiles on some CTAs
    u1, g1 = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse")
    for flags in (9, 0, 1, 8, 4, 12):
        try:
            C.set_option("rg_flags", flags)
            u2, g2 = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse")
            torch.cuda.synchronize()
        finally:
            C.set_option("rg_flags", 28)
        assert torch.equal(u1, u2) and torch.equal(g1, g2), flags


@pytest.mark.parametrize("udf_type,scale", [("square", 1.0), ("sdf", 1.0), ("abs", 0.5)])
def test_reverse_mode_udf_types_and_scale(udf_type, scale):
    from emap_b200 import ops, _cabi as C
    from oracle import emap_oracle as O
    net, p = _net(True, udf_type=udf_type, scale=scale)
    p.udf_type, p.scale = udf_type, scale
    torch.manual_seed(11)
    x = (torch.rand(700, 3) * 2 - 1) * 0.9
    udf, grad = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x.cuda(), mode="reverse")
    ref_u = O.udf_forward(p, x)[0][:, 0]
    ref_g = O.udf_gradient(p, x).detach()
    assert maxdiff(udf.cpu(), ref_u) <= 5e-5 * max(1.0, float(ref_u.abs().max()))
    assert maxdiff(grad.cpu(), ref_g) <= 5e-5 * max(1.0, float(ref_g.abs().max()))


def test_render_with_reverse_mode_matches_default():
    """the drop-in renderer end to end with K1r selected: same outputs as with K1g within the MLP tolerance."""
    from emap_b200 import ops
    from emap_b200.udf_model import BetaNetwork, SingleVarianceNetwork, UDFNetwork
    from emap_b200.udf_renderer_blending import UDFRendererBlending
    from oracle import emap_oracle as O
    torch.manual_seed(0)
    net = UDFNetwork(3, 1, 256, 8, skip_in=[4], multires=10).cuda()
    var, beta = SingleVarianceNetwork(0.3).cuda(), BetaNetwork(0.5, 0.3, 0.3, 5e-5, True, True, False).cuda()
    r = UDFRendererBlending(None, net, var, beta, n_samples=64, n_importance=64, n_outside=0,
                            up_sample_steps=4, perturb=1.0, device="cuda")
    B = 256
    o, d = O.synthetic_rays(B)
    near, far, ds = torch.full((B, 1), 0.05), torch.full((B, 1), 6.0), torch.ones(B, 1)
    outs = {}
    for mode in ("forward", "reverse"):
        ops.set_grad_mode(mode)
        try:
            torch.manual_seed(7)
            with torch.no_grad():
                outs[mode] = r.render(o.cuda(), d.cuda(), near.cuda(), far.cuda(), ds.cuda(),
                                      cos_anneal_ratio=1.0, flip_saturation=0.9)
        finally:
            ops.set_grad_mode(ops.DEFAULT_GRAD_MODE)
    for k in ("edge", "depth", "weights", "gradients", "udf"):
        assert maxdiff(outs["forward"][k], outs["reverse"][k]) <= 2e-3, k


# ---------------------------------------------------------------------------------------------------
# Shared-forward backward (opt-in, ops.set_backward_mode("shared")): K1r writes the value rows of the
# backward's stashes in the training forward, emap_bwd_tangent_forward (mlp_kernel MODE 3) adds the tangent
# rows; everything downstream (top, reverse sweep, dW GEMMs, weight-norm) is the validated path.
# ---------------------------------------------------------------------------------------------------
def test_shared_stash_matches_dual_forward_stash():
    """diagnostic, ops level: (K1r value rows + tangent forward) vs the validated dual forward's stashes.
    fp16 stashes; the value rows now come from the fp32x3 forward instead of a single-fp16-MMA recompute,
    so they agree to the fp16 forward's own error (<= 4e-3 of the plane's max), not bit for bit."""
    from emap_b200 import ops, _cabi as C
    import ctypes
    net, p = _net(True)
    torch.manual_seed(5)
    P = 1000                                           # ragged: 8 tiles of 128, 16 tiles of 64
    x = ((torch.rand(P, 3) * 2 - 1) * 0.9).cuda()
    gbar = torch.randn(P, 3).cuda() * 0.1
    L, desc, st = C.lib(), ctypes.byref(net.desc), C.stream()
    u0_d, u_d = ops.alloc_backward_stash(P, x.device)
    C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), C.PREC_HALF, C.ptr(x), None, None, None, 0, P,
                                    C.ptr(gbar), None, C.ptr(u0_d), C.ptr(u_d), st))
    u0_s, u_s = ops.alloc_backward_stash(P, x.device)
    u0_s.fill_(float("nan")); u_s.fill_(float("nan"))  # every row must be written by one of the two kernels
    ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse", stash=(u0_s, u_s))
    C.check(L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(x), None, None, None, 0, P, C.ptr(gbar),
                                       None, C.ptr(u0_s), C.ptr(u_s), st))
    torch.cuda.synchronize()
    assert torch.isfinite(u0_s).all() and torch.isfinite(u_s).all()
    assert maxdiff(u0_s[:P], u0_d[:P]) <= 1e-3 * float(u0_d[:P].abs().max())   # PE value rows (same sincosf)
    assert maxdiff(u0_s[P:], u0_d[P:]) <= 2e-3 * float(u0_d[P:].abs().max())   # PE tangent rows
    for l in range(8):
        for rows, name in ((slice(0, P), "value"), (slice(P, 2 * P), "tangent")):
            ref = u_d[l][rows].float()
            err = maxdiff(u_s[l][rows].float(), ref) / (float(ref.abs().max()) + 1e-12)
            assert err <= 8e-3, (l, name, err)


@pytest.mark.parametrize("tag,pert", [("init", False), ("pert", True)])
def test_shared_backward_param_grads_vs_reference(golden, tag, pert):
    """the double-backward fixture of tests/test_gpu_train.py with K1r + shared-forward backward selected"""
    from emap_b200 import ops
    from tests.test_gpu_render import build
    from tests.test_gpu_train import _check
    g = golden(f"mlp_{tag}")
    ops.set_grad_mode("reverse"); ops.set_backward_mode("shared")
    try:
        net, var, beta, r = build(10, pert, n_samples=64, n_importance=0, up_sample_steps=5)
        x = g["x"].cuda()
        y, _ = net(x)
        gg = net.gradient(x.clone()).squeeze(1)
        loss = (g["cu"].cuda() * y).sum() + (g["cg"].cuda() * gg).sum()
        assert abs(float(loss) - float(g["loss"])) <= 2e-3 * max(1.0, abs(float(g["loss"])))
        net.zero_grad()
        loss.backward()
        torch.cuda.synchronize()
    finally:
        ops.set_grad_mode(ops.DEFAULT_GRAD_MODE); ops.set_backward_mode(ops.DEFAULT_BWD_MODE)
    _check([(n, p.grad) for n, p in net.named_parameters()], g, "dgrad", 3e-3)     # measured 1.65e-3


def test_shared_backward_render_loss_matches_default():
    """a full render() loss, backward through the drop-in classes: shared-forward vs the default path"""
    from emap_b200 import ops
    from tests.test_gpu_render import build
    from oracle import emap_oracle as O
    B = 300
    o, d = O.synthetic_rays(B)
    near, far, ds = torch.full((B, 1), 0.05), torch.full((B, 1), 6.0), torch.ones(B, 1)
    te = torch.rand(B, 1, generator=torch.Generator().manual_seed(3)).cuda()
    grads = {}
    for mode in (("forward", "dual"), ("reverse", "shared")):
        ops.set_grad_mode(mode[0]); ops.set_backward_mode(mode[1])
        try:
            net, var, beta, r = build(10, True, n_samples=64, n_importance=64, up_sample_steps=4)
            torch.manual_seed(7)
            out = r.render(o.cuda(), d.cuda(), near.cuda(), far.cuda(), ds.cuda(), cos_anneal_ratio=1.0,
                           flip_saturation=0.9)
            loss = (torch.nn.functional.mse_loss(out["edge"], te) + 0.01 * out["gradient_error_near_surface"]
                    + 0.1 * out["gradient_error"])
            loss.backward()
            torch.cuda.synchronize()
            grads[mode] = (float(loss), [p.grad.clone() for p in net.parameters()])
        finally:
            ops.set_grad_mode(ops.DEFAULT_GRAD_MODE); ops.set_backward_mode(ops.DEFAULT_BWD_MODE)
    (l0, g0), (l1, g1) = grads[("forward", "dual")], grads[("reverse", "shared")]
    assert abs(l0 - l1) <= 1e-3 * max(1.0, abs(l0))
    for a, b in zip(g0, g1):
        assert maxdiff(a, b) <= 2e-2 * (float(a.abs().max()) + 1e-12)


def test_rolled_issuer_variants_of_the_backward_are_bit_identical(golden):
    """A/B switches that only change code layout or data movement -- the unrolled MMA-issuer loops of round 1
    (cluster=3 for the K1 family, rg_flags without bits 2 and 4 for K1r), register-staged instead of TMA-staged stash
    rows in the tangent forward (tan_tma=0) and the training forward (rg_flags bit 5) -- must not change a single
    bit of the parameter gradients."""
    from emap_b200 import _cabi as C
    from tests.test_gpu_render import build
    g = golden("mlp_pert")

    def grads():
        net, var, beta, r = build(10, True, n_samples=64, n_importance=0, up_sample_steps=5)
        x = g["x"].cuda().repeat(20, 1)
        y, _ = net(x)
        gg = net.gradient(x.clone()).squeeze(1)
        loss = (g["cu"].cuda().repeat(20, 1) * y).sum() + (g["cg"].cuda().repeat(20, 1) * gg).sum()
        loss.backward()
        torch.cuda.synchronize()
        return [p.grad.clone() for p in net.parameters()]

    ref = grads()
    try:
        C.set_option("cluster", 3); C.set_option("tan_tma", 0); C.set_option("rg_flags", 8 | 32)
        got = grads()
    finally:
        C.set_option("cluster", 1); C.set_option("tan_tma", 1); C.set_option("rg_flags", 28)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
