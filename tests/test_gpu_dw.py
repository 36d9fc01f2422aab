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


@pytest.mark.parametrize("P", [37, 64, 1000, 25600, 25613])
def test_reverse_sweep_tma_staged_is_bit_identical_to_register_staged(P):
    """The default reverse sweep stages both stashes through shared memory with the TMA engine (boxes of 64 points,
    zero fill / clipping past P); the register-staged kernel of round 1 (rev_tma = 0) is the same arithmetic.  Random
    but realistic stashes (h >= 0), every stored row compared bit for bit: fewer points than a tile, exact tiles,
    ragged tails, several tiles per CTA."""
    from emap_b200 import ops, _cabi as C
    torch.manual_seed(P)
    dev = "cuda"
    net = ops.PackedNet(10)
    from tests.helpers import oracle_params
    p = oracle_params(True)
    net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).to(dev))
    st_u = torch.empty(8, 2 * P, 256, device=dev, dtype=torch.float16)
    st_u[:, :P] = (torch.rand(8, P, 256, device=dev) * 0.08).half()          # h: softplus outputs, sigma in (0, 1)
    st_u[:, P:] = (torch.randn(8, P, 256, device=dev) * 0.5).half()           # hdot
    coef = torch.randn(2 * P, device=dev) * 0.3
    L, desc = C.lib(), ctypes.byref(net.desc)
    outs = []
    for tma in (1, 0):
        st_a = torch.full((8, 2 * P, 256), float("nan"), device=dev, dtype=torch.float16)
        try:
            C.set_option("rev_tma", tma)
            C.check(L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(st_u), C.ptr(st_a), P,
                                             C.stream()))
            torch.cuda.synchronize()
        finally:
            C.set_option("rev_tma", 1)
        outs.append(st_a)
    assert torch.isfinite(outs[0].float()).all()                             # every row of every plane was written
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))


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
