"""Timing experiment: tangent forward and reverse sweep with their stash stores / loads switched off (results are
garbage) -- how much of each kernel is its global I/O?  usage: python tools/gpu/gpu_probe_stash_io.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C  # noqa: E402
from tests.helpers import oracle_params  # noqa: E402

p = oracle_params(True)
net = ops.PackedNet(10)
net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).cuda())
P = 1 << 20
x = (torch.rand(P, 3, device="cuda") * 2 - 1) * 1.5
gbar = torch.randn(P, 3, device="cuda") * 1e-3
dudf = torch.randn(P, device="cuda") * 1e-3
L, desc, st = C.lib(), ctypes.byref(net.desc), C.stream()
stash = ops.alloc_backward_stash(P, x.device)
ops.udf_forward_grad(net, 3, pts=x, mode="reverse", stash=stash)
scales = torch.empty(8, device="cuda")
L.emap_bwd_cotangent_scales(C.ptr(dudf), C.ptr(gbar), P, C.ptr(scales), st)
st_a = torch.empty(8, 2 * P, 256, dtype=torch.float16, device="cuda")
coef = torch.randn(2 * P, device="cuda") * 1e-3


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tan = lambda: L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(x), None, None, None, 0, P, C.ptr(gbar),  # noqa: E731
                                         C.ptr(scales), C.ptr(stash[0]), C.ptr(stash[1]), st)
rev = lambda: L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(stash[1]), C.ptr(st_a), P, st)  # noqa: E731
for rnd in range(2):
    C.set_option("tan_tma", 1)
    out = [f"TMA-staged {t(tan):.2f} |"]
    C.set_option("tan_tma", 0)
    for d, name in ((0, "register-staged: full"), (8, "no stores"), (32, "no loads"), (40, "no loads, no stores"), (44, "no epilogue math either")):
        C.set_option("dbg", d)
        out.append(f"{name} {t(tan):.2f}")
    C.set_option("dbg", 0)
    C.set_option("tan_tma", 1)
    print(f"tangent forward round {rnd}: " + "   ".join(out) + "  ms", flush=True)
    out = [f"TMA-staged {t(rev):.2f}  (register-staged form, removed: 5.9-6.3 full, 4.6-5.0 without stores, 4.0-4.5 without loads, 3.0-3.2 without both)"]
    print(f"reverse sweep round {rnd}: " + "   ".join(out) + "  ms", flush=True)
