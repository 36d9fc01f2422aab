import sys, os, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C
from tests.helpers import oracle_params
p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
net = ops.PackedNet(10); net.fold(flat)
P = 1 << 20
x = (torch.rand(P, 3, device="cuda") * 2 - 1) * 1.5
dg = torch.randn(P, 3, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
L = C.lib(); desc = ctypes.byref(net.desc); st = C.stream()
h16 = lambda *s: torch.empty(*s, dtype=torch.float16, device="cuda")
st_u0, st_u = h16(2 * P, 64), h16(8, 2 * P, 256)

def dual(prec): C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), prec, C.ptr(x), None, None, None, 0, P, C.ptr(dg), C.ptr(st_u0), C.ptr(st_u), st))
for dbg in (0, 8, 4, 12, 1, 9):
    C.set_option("dbg", dbg)
    print(f"dbg={dbg} (1 noMMA, 4 noEpiMath, 8 noStashStores): dual fwd 1-term {t(lambda: dual(1)):.2f} ms   3-term {t(lambda: dual(3)):.2f} ms", flush=True)
C.set_option("dbg", 0)
print("fwd1 (128 pts/tile, no stash) for reference:", round(t(lambda: ops.udf_forward(net, 1, pts=x)), 2), "ms")
