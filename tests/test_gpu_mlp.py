"""GPU parity: tcgen05 MLP kernels (K0 fold, K1 forward, K1g forward+gradient) vs the fixtures
generated from the reference and vs the CPU oracle.  Calls go through the C ABI (ctypes).

Tolerances (stated, fp32 reference):
  fp32x3 mode : |udf - ref| <= 5e-5 * max(1,|ref|max);  |grad - ref| <= 5e-5 * max(1,|grad|max)
                measured on B200: 2.2e-5 / 1.4e-5 absolute on udf in [0,2.4], |grad| <= 1.33.  The
                reference's own fp32-CPU-vs-fp64 error on the same inputs is 8e-7; the gap is the tensor
                core's fp32 accumulator, which truncates (round-toward-zero) after every K=16 MMA step
                (48 steps per 256-wide layer in the 3-term scheme) -- see DESIGN.md "precision".
  fp16 mode   : |udf - ref| <= 3e-3 * scale;  |grad - ref| <= 1e-2 * scale   (11-bit operands)
"""
import pytest
import torch

from tests.helpers import maxdiff, oracle_params

pytestmark = pytest.mark.gpu


def _net(pert, multires=10, elem="fp16"):
    from emap_b200 import ops
    p = oracle_params(pert, multires)
    flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
    net = ops.PackedNet(multires, elem_type=elem)
    net.fold(flat)
    return net, p


@pytest.mark.parametrize("pert", [False, True])
def test_wn_fold_matches_torch(pert):
    import ctypes
    import numpy as np
    from oracle import emap_oracle as O
    net, p = _net(pert)
    torch.cuda.synchronize()
    raw = net.packed.cpu().numpy()
    hdr = np.frombuffer(raw[:256].tobytes(), dtype=np.uint32)
    weff_off = hdr[10:19]
    W = O.effective_weights(p)
    for l in range(9):
        n = W[l].numel()
        got = torch.from_numpy(np.frombuffer(raw[weff_off[l]:weff_off[l] + 4 * n].tobytes(),
                                             dtype=np.float32).copy()).reshape(W[l].shape)
        assert maxdiff(got, W[l]) <= 2e-7 * float(W[l].abs().max()) + 1e-12, l


@pytest.mark.parametrize("pert", [False, True])
@pytest.mark.parametrize("prec,tol_u,tol_g", [("fp32", 5e-5, 5e-5), ("fp16", 3e-3, 1e-2)])
def test_forward_and_gradient_vs_reference(golden, pert, prec, tol_u, tol_g):
    from emap_b200 import ops, _cabi as C
    g = golden("mlp_pert" if pert else "mlp_init")
    net, _ = _net(pert)
    x = g["x"].cuda()
    udf, pe = ops.udf_forward(net, C.PRECISIONS[prec], pts=x, want_pe=True)
    udf2, grad = ops.udf_forward_grad(net, C.PRECISIONS[prec], pts=x)
    torch.cuda.synchronize()
    ref_u, ref_g = g["udf"][:, 0], g["grad"][:, 0]
    su = max(1.0, float(ref_u.abs().max()))
    sg = max(1.0, float(ref_g.abs().max()))
    assert maxdiff(pe.cpu(), g["pe"]) <= 5e-7          # sincosf vs torch CPU sin/cos (<= 2 ulp)
    assert maxdiff(udf.cpu(), ref_u) <= tol_u * su
    assert maxdiff(udf2.cpu(), ref_u) <= tol_u * su
    assert maxdiff(grad.cpu(), ref_g) <= tol_g * sg


def test_multires6(golden):
    from emap_b200 import ops, _cabi as C
    g = golden("mlp_mr6_pert")
    net, _ = _net(True, multires=6)
    x = g["x"].cuda()
    udf, grad = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x)
    assert maxdiff(udf.cpu(), g["out"][:, 0]) <= 5e-5 * max(1.0, float(g["out"].abs().max()))
    assert maxdiff(grad.cpu(), g["grad"][:, 0]) <= 5e-5 * max(1.0, float(g["grad"].abs().max()))


def test_ray_addressing_and_ragged_sizes():
    """points given as rays (o + d*z), P not a multiple of the tile, many tiles per CTA."""
    from emap_b200 import ops, _cabi as C
    from oracle import emap_oracle as O
    net, p = _net(True)
    B, n = 1237, 7
    o, d = O.synthetic_rays(B)
    z = torch.rand(B, n) * 3 + 0.5
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).reshape(-1, 3)
    ref = O.udf_forward(p, pts)[0][:, 0]
    udf, _ = ops.udf_forward(net, C.PREC_FP32X3, rays_o=o.cuda(), rays_d=d.cuda(), z=z.cuda())
    u2, gr = ops.udf_forward_grad(net, C.PREC_FP32X3, rays_o=o.cuda(), rays_d=d.cuda(), z=z.cuda())
    refg = O.udf_gradient(p, pts).detach()
    assert maxdiff(udf.cpu(), ref) <= 5e-5 * max(1.0, float(ref.abs().max()))
    assert maxdiff(u2.cpu(), ref) <= 5e-5 * max(1.0, float(ref.abs().max()))
    assert maxdiff(gr.cpu(), refg) <= 5e-5 * max(1.0, float(refg.abs().max()))


@pytest.mark.parametrize("cl", [2, -2, 3])
def test_cluster_weight_stream_variants(golden, cl):
    """cluster=2: multicast pairs (cta_group::1); cluster=-2: CTA pairs driven by one cta_group::2 issuer
    (M=256 MMAs, each CTA stages half of every weight operand); cluster=3: width 1 with the issuer's layer loop
    unrolled (the round-1 form; the default rolls it).  All must be bit-identical to the default."""
    from emap_b200 import ops, _cabi as C
    g = golden("mlp_pert")
    net, _ = _net(True)
    x = g["x"].cuda().repeat(40, 1)           # 15360 points -> 120 / 480 tiles
    C.set_option("cluster", 1)
    u1, g1 = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="forward")
    f1, _ = ops.udf_forward(net, C.PREC_HALF, pts=x)
    try:
        C.set_option("cluster", cl)
        u2, g2 = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="forward")
        f2, _ = ops.udf_forward(net, C.PREC_HALF, pts=x)
        torch.cuda.synchronize()
    finally:
        C.set_option("cluster", 1)
    assert torch.equal(u1, u2) and torch.equal(g1, g2) and torch.equal(f1, f2)


def test_bf16_operands(golden):
    from emap_b200 import ops, _cabi as C
    g = golden("mlp_pert")
    net, _ = _net(True, elem="bf16")
    x = g["x"].cuda()
    udf, grad = ops.udf_forward_grad(net, C.PREC_HALF, pts=x)
    assert maxdiff(udf.cpu(), g["udf"][:, 0]) <= 3e-2
    u3, _ = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x)     # bf16 hi+lo = 16-bit operands
    assert maxdiff(u3.cpu(), g["udf"][:, 0]) <= 2e-4
