import sys, os, torch, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C
from tests.helpers import oracle_params
p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
net = ops.PackedNet(10); net.fold(flat); net.flat = flat; net.fold_id = 1
P = 1 << 20
x = (torch.rand(P, 3, device="cuda") * 2 - 1) * 1.5
du = torch.randn(P, device="cuda"); dg = torch.randn(P, 3, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("full backward (fused):", round(t(lambda: ops.udf_backward(net, 3, du, dg, pts=x, flat_params=flat)), 2), "ms")
os.environ["EMAP_BWD"] = "layerwise"
print("full backward (layerwise):", round(t(lambda: ops.udf_backward(net, 3, du, dg, pts=x, flat_params=flat)), 2), "ms")
del os.environ["EMAP_BWD"]
L = C.lib(); desc = ctypes.byref(net.desc); st = C.stream()
h16 = lambda *s: torch.empty(*s, dtype=torch.float16, device="cuda")
st_u0, st_u = h16(2 * P, 64), h16(8, 2 * P, 256)
st_a = h16(8, 2 * P, 256)
coef = torch.randn(2 * P, device="cuda")
print("dual forward kernel:", round(t(lambda: C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), 1, C.ptr(x), None, None, None, 0, P, C.ptr(dg), C.ptr(st_u0), C.ptr(st_u), st))), 2), "ms")
print("reverse sweep kernel:", round(t(lambda: C.check(L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(st_u), C.ptr(st_a), P, st))), 2), "ms")
A = st_a[3]; U = st_u[2]
print("one dW GEMM [256x2M]@[2Mx256] (torch.mm fp16->fp32):", round(t(lambda: torch.mm(A.t(), U, out_dtype=torch.float32)), 2), "ms")
print("one db sum:", round(t(lambda: A[:P].sum(dim=0, dtype=torch.float32)), 2), "ms")
print("dW via bmm split-K 64:", round(t(lambda: torch.bmm(A.view(64, -1, 256).transpose(1, 2), U.view(64, -1, 256)).float().sum(0)), 2), "ms")
