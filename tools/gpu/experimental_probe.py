"""One opt-in kernel against the validated kernel it would replace: parity on the same inputs + time, as one
JSON line.  Run by bench.py in a child process per item (a trapped launch must not take the bench down) and
reported under the bench line's "experimental" key -- informational, never part of `value` / `e2e`.

usage: python tools/gpu/experimental_probe.py --item k1r|k1_dot|shared_backward|rev2 [--points P]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C  # noqa: E402
from emap_b200.udf_model import UDFNetwork  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--item", required=True, choices=["k1r", "k1_dot", "shared_backward", "rev2"])
    ap.add_argument("--points", type=int, default=1 << 20)
    a = ap.parse_args()
    P = a.points
    # geometric init (the reference's constructor) + a small perturbation so that every PE column matters
    torch.manual_seed(0)
    module = UDFNetwork(d_in=3, d_out=1, d_hidden=256, n_layers=8, skip_in=[4], multires=10, bias=0.5, scale=1.0,
                        geometric_init=True, weight_norm=True, udf_type="abs")
    flat = torch.cat([q.detach().reshape(-1) for q in module.flat_param_list()])
    flat = flat + 0.02 * torch.randn(flat.shape, generator=torch.Generator().manual_seed(1)) * flat.abs().clamp(min=0.05)
    net = ops.PackedNet(10)
    net.fold(flat.cuda())
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.rand(P, 3, device="cuda", generator=g) * 2 - 1) * 1.5
    out = {"item": a.item, "points": P}
    F = 918016.0
    if a.item == "k1r":
        uf, gf = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="forward")
        ur, gr = ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse")
        torch.cuda.synchronize()
        out["max_abs_diff_udf_vs_k1g"] = float((ur - uf).abs().max())
        out["max_abs_diff_grad_vs_k1g"] = float((gr - gf).abs().max())
        out["k1g_ms"] = timed(lambda: ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="forward"))
        out["k1r_ms"] = timed(lambda: ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse"))
        C.set_option("rg_flags", 1)
        out["k1r_split_tail_ms"] = timed(lambda: ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse"))
        C.set_option("rg_flags", 0)
        out["k1r_algorithmic_tflops"] = 2 * F * P / (out["k1r_ms"] * 1e-3) / 1e12
        try:                                                      # extras: must not cost the results above
            uf16, gf16 = ops.udf_forward_grad(net, C.PREC_HALF, pts=x, mode="forward")
            ur16, gr16 = ops.udf_forward_grad(net, C.PREC_HALF, pts=x, mode="reverse")
            torch.cuda.synchronize()
            out["fp16_max_abs_diff_grad_vs_k1g"] = float((gr16 - gf16).abs().max())
            out["fp16_k1g_ms"] = timed(lambda: ops.udf_forward_grad(net, C.PREC_HALF, pts=x, mode="forward"), 3)
            out["fp16_k1r_ms"] = timed(lambda: ops.udf_forward_grad(net, C.PREC_HALF, pts=x, mode="reverse"), 3)
            C.set_option("rg_flags", 2)
            out["k1r_l2_persist_ms"] = timed(lambda: ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse"), 3)
            C.set_option("rg_flags", 0)
        except Exception as e:
            out["extras_error"] = repr(e)[:200]
            C.set_option("rg_flags", 0)
    elif a.item == "k1_dot":
        u0, _ = ops.udf_forward(net, C.PREC_FP32X3, pts=x)
        C.set_option("k1_dot", 1)
        u1, _ = ops.udf_forward(net, C.PREC_FP32X3, pts=x)
        torch.cuda.synchronize()
        out["max_abs_diff_udf_vs_k1"] = float((u1 - u0).abs().max())
        out["k1_dot_ms"] = timed(lambda: ops.udf_forward(net, C.PREC_FP32X3, pts=x))
        C.set_option("k1_dot", 0)
        out["k1_ms"] = timed(lambda: ops.udf_forward(net, C.PREC_FP32X3, pts=x))
    else:
        L, desc, st = C.lib(), ctypes.byref(net.desc), C.stream()
        gbar = torch.randn(P, 3, device="cuda", generator=g) * 0.1
        u0_d, u_d = ops.alloc_backward_stash(P, x.device)

        def dual():
            C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), C.PREC_HALF, C.ptr(x), None, None, None, 0, P,
                                            C.ptr(gbar), C.ptr(u0_d), C.ptr(u_d), st))
        dual()
        if a.item == "shared_backward":
            u0_s, u_s = ops.alloc_backward_stash(P, x.device)
            ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse", stash=(u0_s, u_s))

            def tangent():
                C.check(L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(x), None, None, None, 0, P,
                                                   C.ptr(gbar), C.ptr(u0_s), C.ptr(u_s), st))
            tangent()
            torch.cuda.synchronize()
            rel = []
            for l in range(8):
                for rows in (slice(0, P), slice(P, 2 * P)):
                    ref = u_d[l][rows].float()
                    rel.append(float((u_s[l][rows].float() - ref).abs().max() / (ref.abs().max() + 1e-12)))
            out["max_rel_diff_stash_vs_dual_forward"] = max(rel)
            out["dual_forward_ms"] = timed(dual, 3)
            out["tangent_forward_ms"] = timed(tangent, 3)
            out["k1r_with_stash_ms"] = timed(
                lambda: ops.udf_forward_grad(net, C.PREC_FP32X3, pts=x, mode="reverse", stash=(u0_s, u_s)), 3)
        else:
            coef = torch.randn(2 * P, device="cuda", generator=g) * 0.5
            res = []
            for tiles in (1, 2):
                st_a = torch.empty(8, 2 * P, 256, dtype=torch.float16, device="cuda")
                C.set_option("rev_tiles", tiles)

                def sweep():
                    C.check(L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(u_d), C.ptr(st_a), P, st))
                out["rev_ms" if tiles == 1 else "rev2_ms"] = timed(sweep, 3)
                res.append(st_a)
            C.set_option("rev_tiles", 1)
            out["rev2_bit_identical"] = bool(torch.equal(res[0], res[1]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
