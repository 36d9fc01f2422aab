"""CPU: every C-ABI call the host shim makes is checked against the declared ctypes signature.

The GPU-only code paths of emap_b200/ops.py cannot execute here, and a call with one argument too many or
a pointer where an int belongs would only surface on the B200 box.  This test swaps the library for a
recorder that validates each call against `_cabi.SIGNATURES` (argument count + ctypes conversion of every
argument) and returns success without touching memory, lets CPU tensors through `_cabi.ptr`, and then
drives the shim's entry points -- including the opt-in K1r / shared-forward-backward paths -- end to end.
Nothing is computed: outputs are whatever torch.empty returned.  (Host-only library functions -- sizes,
parameter counts -- are forwarded to the real library.)
"""
import ctypes

import pytest
import torch

from emap_b200 import _cabi as C

HOST_ONLY = {"emap_flat_param_count", "emap_packed_size", "emap_last_error", "emap_abi_version",
             "emap_packed_offsets", "emap_set_option"}


class _Recorder:
    def __init__(self, real):
        self.real = real
        self.calls = []

    def __getattr__(self, name):
        if name not in C.SIGNATURES:
            raise AttributeError(f"{name} is not declared in _cabi.SIGNATURES")
        restype, argtypes = C.SIGNATURES[name]
        if name in HOST_ONLY:
            return getattr(self.real, name)

        def call(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} arguments, signature has {len(argtypes)}"
            for i, (a, t) in enumerate(zip(args, argtypes)):
                try:
                    t.from_param(a)
                except (TypeError, ctypes.ArgumentError) as e:     # pragma: no cover - failure path
                    raise AssertionError(f"{name}: argument {i} ({a!r}) does not convert to {t}: {e}")
            self.calls.append(name)
            if name == "emap_rgrad_scratch_bytes":
                return 4 * 7 * 4 * 16 * 512 * 4                    # pretend 4 SMs
            if name == "emap_bwd_workspace_bytes":
                return 1 << 16
            if name == "emap_bwd_weight_grads":
                return 4                                           # number of partials
            return 0
        return call


@pytest.fixture()
def shim(monkeypatch):
    from emap_b200 import ops
    rec = _Recorder(C.lib())
    monkeypatch.setattr(C, "lib", lambda: rec)
    monkeypatch.setattr(C, "ptr", lambda t: None if t is None else (t.contiguous().data_ptr() or 1))
    monkeypatch.setattr(C, "stream", lambda: 0)
    monkeypatch.setattr(ops, "_rg_scratch", lambda dev: torch.empty(1 << 20, dtype=torch.uint8))
    # library GEMMs of the backward: fp16 x fp16 -> fp32 is a CUDA feature; shapes are what matters here
    real_mm = torch.mm
    monkeypatch.setattr(torch, "mm", lambda a, b, out_dtype=None: real_mm(a.float(), b.float()))
    yield ops, rec
    ops.set_grad_mode(ops.DEFAULT_GRAD_MODE)
    ops.set_backward_mode(ops.DEFAULT_BWD_MODE)


def _packed(ops, multires=10):
    net = ops.PackedNet(multires, device="cpu")
    net.fold(torch.zeros(net.n_params))
    net.flat = torch.zeros(net.n_params)
    return net


def test_forward_entry_points(shim):
    ops, rec = shim
    net = _packed(ops)
    x = torch.zeros(37, 3)
    o, d, z = torch.zeros(5, 3), torch.zeros(5, 3), torch.zeros(5, 7)
    ops.udf_forward(net, C.PREC_FP32X3, pts=x, want_pe=True)
    ops.udf_forward(net, C.PREC_HALF, rays_o=o, rays_d=d, z=z)
    for mode in ("forward", "reverse"):
        u, g = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode=mode)
        assert u.shape == (37,) and g.shape == (37, 3)
        ops.udf_forward_grad(net, C.PREC_FP32X3, rays_o=o, rays_d=d, z=z, mode=mode)
    stash = ops.alloc_backward_stash(37, "cpu")
    ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse", stash=stash)
    with pytest.raises(RuntimeError):
        ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="forward", stash=stash)    # only K1r fills a stash
    with pytest.raises(RuntimeError):
        ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse", stash=ops.alloc_backward_stash(36, "cpu"))
    ops.debug_mlp(net, C.PREC_FP32X3, 1, x)
    ops.debug_rgrad(net, C.PREC_FP32X3, x)
    for name in ("emap_wn_fold", "emap_udf_forward", "emap_udf_forward_grad", "emap_udf_forward_grad_rev",
                 "emap_debug_mlp", "emap_debug_rgrad"):
        assert name in rec.calls, name


@pytest.mark.parametrize("shared", [False, True])
def test_backward_entry_points(shim, shared):
    ops, rec = shim
    net = _packed(ops)
    P = 24
    x = torch.zeros(P, 3)
    stash = ops.alloc_backward_stash(P, "cpu") if shared else None
    flat_grad = ops.udf_backward(net, C.PREC_FP32X3, torch.zeros(P), torch.zeros(P, 3), pts=x,
                                 flat_params=net.flat, stash=stash)
    assert flat_grad.shape == net.flat.shape
    assert ("emap_bwd_tangent_forward" in rec.calls) == shared
    assert ("emap_bwd_dual_forward" in rec.calls) == (not shared)
    for name in ("emap_bwd_cotangent_scales", "emap_bwd_top", "emap_bwd_reverse_sweep", "emap_bwd_weight_grads",
                 "emap_bwd_finish"):
        assert name in rec.calls, name


def test_mode_switches_validate(shim):
    ops, _ = shim
    with pytest.raises(ValueError):
        ops.set_grad_mode("sideways")
    with pytest.raises(ValueError):
        ops.set_backward_mode("borrowed")
    ops.set_grad_mode("forward")
    ops.set_backward_mode("shared")
    assert not ops.shared_backward()            # needs the reverse-mode forward
    ops.set_grad_mode("reverse")
    assert ops.shared_backward()


def test_ray_kernels_and_callers(shim):
    ops, rec = shim
    B, n, k = 6, 16, 4
    z = torch.zeros(B, n)
    o = d = torch.zeros(B, 3)
    lin = torch.linspace(0, 1, n)
    ops.coarse_z(torch.zeros(1), torch.ones(1), False, lin, torch.zeros(B), B, n)
    sd = torch.zeros(1)
    u = torch.linspace(0.1, 0.9, k)
    ops.upsample_step(o, d, z, torch.zeros(B, n), None, None, u, k, sd, 64.0, 128.0, 20.0, want_inds=True,
                      want_weights=True)
    ops.upsample_step(o, d, z, torch.zeros(B, n), torch.zeros(B, k), torch.zeros(B, k), u, k, sd, 64.0, 128.0, 20.0)
    ops.sample_pdf_det(z, torch.zeros(B, n - 1), k)
    ops.render_prep(z, sd)
    ops.null_direction(torch.zeros(5, 50, 3))
    for name in ("emap_coarse_z", "emap_upsample_step", "emap_render_prep", "emap_null_direction"):
        assert name in rec.calls, name


class _FakeModule:
    """just enough of emap_b200.UDFNetwork for the autograd Functions: a CPU PackedNet behind the recorder"""

    def __init__(self, ops):
        self.net = _packed(ops)
        self.params = [torch.zeros(n, requires_grad=True) for n in (self.net.n_params - 10, 10)]
        self.prec_code = C.PREC_FP32X3
        self.net.fold_id = 1

    def packed(self):
        return self.net

    def flat_param_list(self):
        return self.params


@pytest.mark.parametrize("grad_mode,bwd_mode,expect_shared", [("forward", "dual", False), ("reverse", "dual", False),
                                                              ("reverse", "shared", True), ("forward", "shared", False)])
def test_autograd_wiring_of_the_backward_modes(shim, grad_mode, bwd_mode, expect_shared):
    """_UDFForwardGrad end to end on the recorder: which forward / stage-1 kernels each mode combination calls,
    that the stash made in forward reaches backward, and that every parameter gets a gradient of its shape."""
    ops, rec = shim
    from emap_b200.autograd import udf_forward_grad_fn
    ops.set_grad_mode(grad_mode)
    ops.set_backward_mode(bwd_mode)
    mod = _FakeModule(ops)
    x = torch.zeros(40, 3)
    udf, grad = udf_forward_grad_fn(mod, x)
    assert ("emap_udf_forward_grad_rev" in rec.calls) == (grad_mode == "reverse")
    (udf.sum() + grad.sum()).backward()
    assert ("emap_bwd_tangent_forward" in rec.calls) == expect_shared
    assert ("emap_bwd_dual_forward" in rec.calls) == (not expect_shared)
    for p in mod.params:
        assert p.grad is not None and p.grad.shape == p.shape
    # parameters re-folded between forward and backward: the stash is stale -> fall back to the dual forward
    if expect_shared:
        rec.calls.clear()
        udf, grad = udf_forward_grad_fn(mod, x)
        mod.net.fold_id += 1
        (udf.sum() + grad.sum()).backward()
        assert "emap_bwd_dual_forward" in rec.calls and "emap_bwd_tangent_forward" not in rec.calls
    # no parameter gradient requested (inference under no_grad) -> no stash is allocated / filled;
    # with grad mode on it is, exactly when the shared backward is selected
    allocs = []
    real_alloc = ops.alloc_backward_stash
    ops.alloc_backward_stash = lambda *a, **k: (allocs.append(a), real_alloc(*a, **k))[1]
    try:
        rec.calls.clear()
        with torch.no_grad():
            udf_forward_grad_fn(mod, x)
        assert rec.calls.count("emap_udf_forward_grad_rev") + rec.calls.count("emap_udf_forward_grad") == 1
        assert allocs == []
        udf_forward_grad_fn(mod, x)
        assert (len(allocs) == 1) == expect_shared
    finally:
        ops.alloc_backward_stash = real_alloc


class _FakeNet(_FakeModule):
    """+ the renderer-facing method of UDFNetwork"""

    def udf_and_gradient(self, x=None, rays_o=None, rays_d=None, z=None):
        from emap_b200.autograd import udf_forward_grad_fn
        return udf_forward_grad_fn(self, x, rays_o, rays_d, z)


def _cpu_renderer(ops, **kw):
    from emap_b200.udf_model import BetaNetwork, SingleVarianceNetwork
    from emap_b200.udf_renderer_blending import UDFRendererBlending
    net = _FakeNet(ops)
    var, beta = SingleVarianceNetwork(0.3), BetaNetwork(0.5, 0.3, 0.3, 5e-5, True, True, False)
    cfg = dict(n_samples=16, n_importance=8, n_outside=0, up_sample_steps=4, perturb=1.0, device="cpu")
    cfg.update(kw)
    return net, UDFRendererBlending(None, net, var, beta, **cfg)


def _no_device_queries(monkeypatch, ops):
    """the status poll and the capture query talk to the CUDA driver: not on this machine"""
    monkeypatch.setattr(ops, "poll_status", lambda dev: None)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)


def test_render_call_sequence_on_the_recorder(shim, monkeypatch):
    """UDFRendererBlending.render() end to end on the recorder library (nothing computes): which C-ABI calls one
    render() makes, in inference and in training, and that nothing but the declared entry points is used."""
    ops, rec = shim
    _no_device_queries(monkeypatch, ops)
    ops.set_grad_mode("reverse"); ops.set_backward_mode("shared")
    net, r = _cpu_renderer(ops)
    B = 5
    o = d = torch.zeros(B, 3)
    rec.calls.clear()
    with torch.no_grad():
        out = r.render(o, d, 0.05, 6.0, torch.ones(B, 1), cos_anneal_ratio=1.0, flip_saturation=0.9)
    assert out["edge"].shape == (B, 1) and out["weights"].shape == (B, 24)
    assert rec.calls.count("emap_coarse_z") == 1
    assert rec.calls.count("emap_udf_forward") == 4                  # coarse samples + 3 of the 4 up-sampling rounds
    assert rec.calls.count("emap_upsample_step") == 5                # 4 rounds + the final merge
    assert rec.calls.count("emap_render_prep") == 1 and rec.calls.count("emap_udf_forward_grad_rev") == 1
    assert rec.calls.count("emap_render_core_fwd") == 1
    assert not any(c.startswith("emap_bwd") for c in rec.calls)
    # training: the same forward, then the backward's stages exactly once each
    rec.calls.clear()
    out = r.render(o, d, 0.05, 6.0, torch.ones(B, 1), cos_anneal_ratio=1.0, flip_saturation=0.9)
    (out["edge"].sum() + out["gradient_error"]).backward()
    for name in ("emap_render_core_bwd", "emap_bwd_cotangent_scales", "emap_bwd_tangent_forward", "emap_bwd_top",
                 "emap_bwd_reverse_sweep", "emap_bwd_weight_grads", "emap_bwd_finish"):
        assert rec.calls.count(name) == 1, (name, rec.calls)
    assert "emap_bwd_dual_forward" not in rec.calls
    for p in net.params:
        assert p.grad is not None


def test_render_refuses_host_draws_while_a_graph_is_prepared(shim, monkeypatch):
    """graph.GraphedStep sets ops.graph_prepare during its warm-up: a host-side stratified draw (perturb > 0 without
    perturb_on_device) is refused there already, before any capture has begun"""
    ops, rec = shim
    _no_device_queries(monkeypatch, ops)
    net, r = _cpu_renderer(ops)
    o = d = torch.zeros(3, 3)
    ops.graph_prepare = True
    try:
        with pytest.raises(RuntimeError, match="perturb_on_device"):
            r.render(o, d, 0.05, 6.0, torch.ones(3, 1), cos_anneal_ratio=1.0)
        r.perturb_on_device = True
        with torch.no_grad():
            r.render(o, d, 0.05, 6.0, torch.ones(3, 1), cos_anneal_ratio=1.0)
    finally:
        ops.graph_prepare = False
