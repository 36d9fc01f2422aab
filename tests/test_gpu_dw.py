"""GPU: the weight-gradient contraction kernel (emap_b200/csrc/mlp_dw.cu) in isolation -- random fp16 stashes,
per-CTA partials read back from the workspace and summed, against fp32 matmuls of the same fp16 data.  Row counts
that are not multiples of the 64-row stage (TMA zero fill), fewer stages than SMs, and value/tangent boundaries
inside a stage (bias sums count value rows only)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

PART = 2 * 256 * 64 + 7 * 256 * 256
# (job offset, n, A plane, U source) in the order of mlp_dw.cu: c_jobs
JOBS = [(0, 64, 0, "u0"), (16384, 256, 1, 0), (16384 + 65536, 256, 2, 1), (16384 + 2 * 65536, 256, 3, 2),
        (16384 + 3 * 65536, 256, 4, 3), (16384 + 4 * 65536, 64, 4, "u0"),
        (32768 + 4 * 65536, 256, 5, 4), (32768 + 5 * 65536, 256, 6, 5), (32768 + 6 * 65536, 256, 7, 6)]


@pytest.mark.parametrize("P", [1000, 70, 4096 * 3 + 17])
def test_weight_grad_partials_match_matmul(P):
    from emap_b200 import ops, _cabi as C
    torch.manual_seed(P)
    dev = "cuda"
    st_a = (torch.randn(8, 2 * P, 256, device=dev) * 0.5).half()
    st_u = (torch.randn(8, 2 * P, 256, device=dev) * 0.5).half()
    st_u0 = (torch.randn(2 * P, 64, device=dev) * 0.5).half()
    net = ops.PackedNet(10)
    ws = ops._bwd_workspace(torch.device(dev))
    ws.zero_()
    n_parts = int(C.lib().emap_bwd_weight_grads(ctypes.byref(net.desc), C.ptr(st_a), C.ptr(st_u0), C.ptr(st_u), P,
                                                C.ptr(ws), ws.numel(), C.stream()))
    torch.cuda.synchronize()
    assert n_parts > 0, C.lib().emap_last_error()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    wsf = ws.view(torch.float32)
    parts = wsf[:sms * PART].view(sms, PART)[:n_parts].double().sum(0)
    dbp = wsf[sms * PART:sms * PART + sms * 2048].view(sms, 8, 256)[:n_parts].double().sum(0)
    for off, n, la, lu in JOBS:
        A = st_a[la].double()
        U = (st_u0 if lu == "u0" else st_u[lu]).double()
        ref = A.t() @ U                                               # [256, n]
        got = parts[off:off + 256 * n].view(256, n)
        err = float((got - ref).abs().max()) / float(ref.abs().max())
        assert err <= 1e-5, (off, n, la, lu, err)                     # fp32 accumulation of exact fp16 products
    for l in range(8):
        ref = st_a[l][:P].double().sum(0)
        assert float((dbp[l] - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), l


def _reverse_sweep_reference(net, st_u, coef, P):
    """fp32 torch restatement of mlp_rev.cu on the same fp16 data: A_7 from the output-layer pull-back, then
    [eta ; etadot] = fp16(A_l) . fp16(16 W_l) / 16  (1/sqrt2 on the skip layer, whose PE inputs are not hidden
    units), alpha = etadot 100 hdot (1 - sigma) + eta sigma, alphadot = etadot sigma, sigma = 1 - exp(-100 h)."""
    from emap_b200 import ops
    W, in_dim, out_dim = ops._weff_views(net)
    pe = 3 + 6 * net.multires
    out3 = 256 - pe
    w8 = W[8].reshape(-1).float()
    eta = coef[:P, None] * w8[None, :]
    etad = coef[P:, None] * w8[None, :]
    planes = []
    for lt in range(7, -1, -1):
        h, hd = st_u[lt, :P].float(), st_u[lt, P:].float()
        one_m_s = torch.exp(-100.0 * h)
        sg = 1.0 - one_m_s
        al = etad * (100.0 * hd * one_m_s) + eta * sg
        ad = etad * sg
        if lt == 3:
            al[:, out3:] = 0.0
            ad[:, out3:] = 0.0
        a16 = torch.cat([al, ad]).half()
        planes.append(a16)
        if lt >= 1:
            mul = 0.70710678118654752440 if lt == 4 else 1.0
            Wl = (W[lt].float() * (mul * 16.0)).half().float() / 16.0          # the kernel's fp16 operand image
            if lt == 4:
                Wl = Wl.clone()
                Wl[:, out3:] = 0.0                                             # PE inputs of the skip layer
            prod = a16[:, :Wl.shape[0]].float() @ Wl                           # [2P, in]  (layer 3 has 256 - pe outputs)
            eta, etad = prod[:P, :256], prod[P:, :256]
    return torch.stack(planes[::-1])                                           # [8, 2P, 256]


@pytest.mark.parametrize("P", [37, 64, 1000, 25600, 25613])
def test_reverse_sweep_matches_fp32_reference(P):
    """The reverse sweep (both stashes staged through shared memory by TMA: boxes of 64 points, zero fill / clipping
    past P) against an fp32 restatement on the same fp16 stashes: fewer points than a tile, exact tiles, ragged
    tails, several tiles per CTA.  Tolerance = fp16 rounding of the stored adjoints (2^-11 relative) compounded
    over the eight layers + accumulation order.  (The kernel replaced a register-staged one of identical arithmetic;
    the two were bit-identical on hardware for these sizes -- profiles/r02_stash_io_probe.txt -- before that one
    was removed.)"""
    from emap_b200 import ops, _cabi as C
    from tests.helpers import oracle_params
    torch.manual_seed(P)
    dev = "cuda"
    net = ops.PackedNet(10)
    p = oracle_params(True)
    net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).to(dev))
    st_u = torch.empty(8, 2 * P, 256, device=dev, dtype=torch.float16)
    st_u[:, :P] = (torch.rand(8, P, 256, device=dev) * 0.08).half()          # h: softplus outputs, sigma in (0, 1)
    st_u[:, P:] = (torch.randn(8, P, 256, device=dev) * 0.5).half()           # hdot
    coef = torch.randn(2 * P, device=dev) * 0.3
    st_a = torch.full((8, 2 * P, 256), float("nan"), device=dev, dtype=torch.float16)
    C.check(C.lib().emap_bwd_reverse_sweep(ctypes.byref(net.desc), C.ptr(net.packed), C.ptr(coef), C.ptr(st_u),
                                           C.ptr(st_a), P, C.stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(st_a.float()).all()                                # every row of every plane was written
    ref = _reverse_sweep_reference(net, st_u, coef, P)
    for lt in range(8):
        scale = float(ref[lt].float().abs().max())
        err = float((st_a[lt].float() - ref[lt].float()).abs().max())
        assert err <= 4e-3 * scale, (lt, err, scale)


@pytest.mark.parametrize("P", [50, 128, 1000, 51200, 51277])
def test_tangent_forward_tma_staged_is_bit_identical_to_register_staged(P):
    """The tangent forward moves the stash rows (value rows in, tangent rows out) through shared memory with the TMA
    engine by default; tan_tma = 0 is the register-staged form of the same arithmetic.  Real value rows (written by the
    training forward K1r), random cotangent direction; the tangent rows of all eight planes and of the PE stash are
    compared bit for bit; the value rows must be untouched."""
    from emap_b200 import ops, _cabi as C
    from tests.helpers import oracle_params
    torch.manual_seed(P)
    dev = torch.device("cuda")
    p = oracle_params(True)
    net = ops.PackedNet(10)
    net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).to(dev))
    x = (torch.rand(P, 3, device=dev) * 2 - 1) * 1.2
    gbar = torch.randn(P, 3, device=dev) * 1e-4
    L, desc, st = C.lib(), ctypes.byref(net.desc), C.stream()
    scales = torch.empty(8, device=dev)
    C.check(L.emap_bwd_cotangent_scales(None, C.ptr(gbar), P, C.ptr(scales), st))
    outs = []
    for tma in (1, 0):
        stash = ops.alloc_backward_stash(P, dev)
        stash[0].fill_(float("nan")); stash[1].fill_(float("nan"))
        ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse", stash=stash)
        values = stash[1][:, :P].clone()
        try:
            C.set_option("tan_tma", tma)
            C.check(L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(x), None, None, None, 0, P, C.ptr(gbar),
                                               C.ptr(scales), C.ptr(stash[0]), C.ptr(stash[1]), st))
            torch.cuda.synchronize()
        finally:
            C.set_option("tan_tma", 1)
        assert torch.equal(stash[1][:, :P].view(torch.int16), values.view(torch.int16))
        outs.append((stash[0].clone(), stash[1].clone()))
    assert torch.isfinite(outs[0][1].float()).all() and torch.isfinite(outs[0][0].float()).all()
    assert torch.equal(outs[0][0].view(torch.int16), outs[1][0].view(torch.int16))
    assert torch.equal(outs[0][1].view(torch.int16), outs[1][1].view(torch.int16))


@pytest.mark.parametrize("P", [90, 128, 3000, 38400, 38411])
def test_training_stash_by_tma_is_bit_identical_to_register_stores(P):
    """K1r in training mode writes the value rows of the backward's stash: by TMA stores from the A tile (default,
    layers 0..6) or with per-thread register stores (rg_flags bit 5, the round-2 form).  Same bits, same udf /
    gradient; rows past P and the tangent half of every plane stay untouched."""
    from emap_b200 import ops, _cabi as C
    from tests.helpers import oracle_params
    torch.manual_seed(P)
    dev = torch.device("cuda")
    p = oracle_params(True)
    net = ops.PackedNet(10)
    net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).to(dev))
    x = (torch.rand(P, 3, device=dev) * 2 - 1) * 1.2
    outs = []
    for flags in (28, 28 | 32):
        stash = ops.alloc_backward_stash(P, dev)
        stash[0].fill_(7.0); stash[1].fill_(7.0)
        try:
            C.set_option("rg_flags", flags)
            u, g = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse", stash=stash)
            torch.cuda.synchronize()
        finally:
            C.set_option("rg_flags", 28)
        assert bool((stash[1][:, P:] == 7.0).all()) and bool((stash[0][P:] == 7.0).all())
        outs.append((u, g, stash[0].clone(), stash[1].clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert not bool((outs[0][3][:, :P] == 7.0).all(dim=-1).any())          # every value row was written
