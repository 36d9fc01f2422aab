"""reverse sweep: default kernel vs two tiles in flight (rev_tiles=2), 1 M points.
usage: python tools/gpu/gpu_time_rev2.py"""
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
gbar = torch.randn(P, 3, device="cuda") * 0.1
L, desc, st = C.lib(), ctypes.byref(net.desc), C.stream()
st_u0, st_u = ops.alloc_backward_stash(P, x.device)
C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), C.PREC_HALF, C.ptr(x), None, None, None, 0, P,
                                C.ptr(gbar), C.ptr(st_u0), C.ptr(st_u), st))
coef = torch.randn(2 * P, device="cuda") * 0.5
st_a = torch.empty(8, 2 * P, 256, dtype=torch.float16, device="cuda")


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


def sweep():
    C.check(L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(st_u), C.ptr(st_a), P, st))


for tiles in (1, 2):
    C.set_option("rev_tiles", tiles)
    print(f"reverse sweep, rev_tiles={tiles}: {t(sweep):.2f} ms", flush=True)
C.set_option("rev_tiles", 1)
