import sys, os, torch
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
st_a = torch.randn(8, 2 * P, 256, device="cuda", dtype=torch.float16)
db = torch.empty(8, 256, device="cuda"); part = torch.empty(8 * 296 * 256, device="cuda")
L = C.lib()
print("bias sums kernel:", round(t(lambda: C.check(L.emap_bwd_bias_sums(C.ptr(st_a), P, C.ptr(part), C.ptr(db), C.stream()))), 3), "ms")
ref = st_a[:, :P].sum(dim=1, dtype=torch.float32)
print("torch reduce:", round(t(lambda: st_a[:, :P].sum(dim=1, dtype=torch.float32)), 3), "ms", " maxdiff", float((db - ref).abs().max()), "scale", float(ref.abs().max()))
