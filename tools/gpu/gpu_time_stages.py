"""Per-stage timing of the MLP kernels of one training step at 1 M points (the bench workload's core points), with
the dynamic tile scheduler on and off, plus bit-identity of the two schedules.
usage: python tools/gpu/gpu_time_stages.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C  # noqa: E402
from tests.helpers import oracle_params  # noqa: E402

p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
net = ops.PackedNet(10)
net.fold(flat)
P = 1 << 20
x = (torch.rand(P, 3, device="cuda") * 2 - 1) * 1.5
gbar = torch.randn(P, 3, device="cuda") * 1e-3
dudf = torch.randn(P, device="cuda") * 1e-3
L, desc, st = C.lib(), ctypes.byref(net.desc), C.stream()


def t(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


stash = ops.alloc_backward_stash(P, x.device)
st_a = torch.empty(8, 2 * P, 256, dtype=torch.float16, device="cuda")
coef = torch.empty(2 * P, device="cuda")
ws = ops._bwd_workspace(x.device)
scales = torch.empty(8, device="cuda")
W, _, _ = ops._weff_views(net)
res = {}
for dyn in (1, 0):
    C.set_option("dynamic_tiles", dyn)
    C.set_option("rg_flags", 28 if dyn else 20)
    tag = "dynamic" if dyn else "static "
    fwd = t(lambda: ops.udf_forward(net, 3, pts=x))
    u0, _ = ops.udf_forward(net, 3, pts=x)
    k1r = t(lambda: ops.udf_forward_grad(net, 3, pts=x, mode="reverse"))
    k1rs = t(lambda: ops.udf_forward_grad(net, 3, pts=x, mode="reverse", stash=stash))
    L.emap_bwd_cotangent_scales(C.ptr(dudf), C.ptr(gbar), P, C.ptr(scales), st)
    tan = t(lambda: L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(x), None, None, None, 0, P, C.ptr(gbar),
                                               C.ptr(scales), C.ptr(stash[0]), C.ptr(stash[1]), st))
    top = t(lambda: L.emap_bwd_top(desc, C.ptr(stash[1][7]), C.ptr(W[8].reshape(-1)), C.ptr(flat[-1:]), C.ptr(dudf),
                                   C.ptr(scales), P, C.ptr(coef), C.ptr(ws), st))
    rev = t(lambda: L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(stash[1]), C.ptr(st_a), P, st))
    dw = t(lambda: L.emap_bwd_weight_grads(desc, C.ptr(st_a), C.ptr(stash[0]), C.ptr(stash[1]), P, C.ptr(ws), ws.numel(), st))
    dual = t(lambda: L.emap_bwd_dual_forward(desc, C.ptr(net.packed), C.PREC_HALF, C.ptr(x), None, None, None, 0, P,
                                             C.ptr(gbar), C.ptr(scales), C.ptr(stash[0]), C.ptr(stash[1]), st))
    torch.cuda.synchronize()
    res[dyn] = (u0.clone(), st_a[:, :4096].clone(), stash[1][:, P:P + 4096].clone())
    print(f"{tag}: K1 forward {fwd:.2f}  K1r {k1r:.2f}  K1r+stash {k1rs:.2f}  tangent fwd {tan:.2f}  top {top:.2f}  "
          f"reverse sweep {rev:.2f}  weight grads {dw:.2f}  (dual forward {dual:.2f})  ms", flush=True)
C.set_option("dynamic_tiles", 1)
C.set_option("rg_flags", 28)
print("bit-identical static vs dynamic:", [bool(torch.equal(a, b)) for a, b in zip(res[0], res[1])])
# code-layout variants (rolled MMA-issuer loops), two interleaved rounds against order effects.
# rg_flags: 8 = dynamic tiles, +4 = rolled issuer, +16 = rolled reverse-epilogue chunk loop (default 28), +2 = persisting L2,
# +32 = training stash from registers instead of by TMA;
# cluster 3 = the unrolled issuer loop of round 1
for rnd in range(2):
    line = []
    for fl in (28, 28 + 32, 12, 30):
        C.set_option("rg_flags", fl)
        line.append(f"flags {fl}: {t(lambda: ops.udf_forward_grad(net, 3, pts=x, mode='reverse', stash=stash), reps=6):.2f}")
    C.set_option("rg_flags", 28)
    print(f"K1r+stash round {rnd}: " + "   ".join(line) + "  ms", flush=True)
    line = []
    for fl in (28, 12, 8):
        C.set_option("rg_flags", fl)
        line.append(f"flags {fl}: {t(lambda: ops.udf_forward_grad(net, 3, pts=x, mode='reverse'), reps=6):.2f}")
    C.set_option("rg_flags", 28)
    print(f"K1r (inference) round {rnd}: " + "   ".join(line) + "  ms", flush=True)
    line = []
    for cl in (1, 3):
        C.set_option("cluster", cl)
        f = t(lambda: ops.udf_forward(net, 3, pts=x), reps=6)
        tn = t(lambda: L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(x), None, None, None, 0, P, C.ptr(gbar),
                                                  C.ptr(scales), C.ptr(stash[0]), C.ptr(stash[1]), st), reps=6)
        line.append(f"cluster {cl}: K1 forward {f:.2f}  tangent fwd {tn:.2f}")
    C.set_option("cluster", 1)
    print(f"issuer variants round {rnd}: " + "   ".join(line) + "  ms", flush=True)
