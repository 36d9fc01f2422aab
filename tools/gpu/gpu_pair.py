"""cta_group::2 pair mode vs the default: bit-identity of every MLP mode, then timing."""
import sys, os, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C
from tests.helpers import oracle_params
p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
net = ops.PackedNet(10); net.fold(flat)
L = C.lib(); desc = ctypes.byref(net.desc); st = C.stream()
h16 = lambda *s: torch.zeros(*s, dtype=torch.float16, device="cuda")

def run_all(P, prec):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.rand(P, 3, device="cuda", generator=g) * 2 - 1) * 1.5
    dg = torch.randn(P, 3, device="cuda", generator=g)
    u0, _ = ops.udf_forward(net, prec, pts=x)
    u1, g1 = ops.udf_forward_grad(net, prec, pts=x)
    st_u0, st_u = h16(2 * P, 64), h16(8, 2 * P, 256)
    C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), prec, C.ptr(x), None, None, None, 0, P, C.ptr(dg),
                                    C.ptr(st_u0), C.ptr(st_u), st))
    torch.cuda.synchronize()
    return u0, u1, g1, st_u0, st_u

ok = True
for prec in (3, 1):
    for P in (64, 1000, 148 * 128 * 2 + 77):
        C.set_option("cluster", 1); ref = run_all(P, prec)
        C.set_option("cluster", -2); out = run_all(P, prec)
        same = [bool(torch.equal(a, b)) for a, b in zip(ref, out)]
        err = [float((a.float() - b.float()).abs().max()) for a, b in zip(ref, out)]
        print(f"prec={prec} P={P}: identical {same}  maxdiff {err}", flush=True)
        ok = ok and all(same)
print("PAIR PARITY", "OK" if ok else "FAILED", flush=True)

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
P = 1 << 20
x = (torch.rand(P, 3, device="cuda") * 2 - 1) * 1.5
for cl in (1, -2):
    C.set_option("cluster", cl)
    r = []
    for prec in (3, 1):
        r.append(t(lambda: ops.udf_forward(net, prec, pts=x)))
        r.append(t(lambda: ops.udf_forward_grad(net, prec, pts=x)))
    print(f"cluster={cl}  fwd3 {r[0]:.2f} ms  grad3 {r[1]:.2f} ms  fwd1 {r[2]:.2f} ms  grad1 {r[3]:.2f} ms", flush=True)
C.set_option("cluster", 1)
