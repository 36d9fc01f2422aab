"""timing experiments on the MLP kernels.  usage: gpu_time_mlp.py"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C
from tests.helpers import oracle_params
p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
net = ops.PackedNet(10); net.fold(flat)
P = 1 << 20
x = (torch.rand(P, 3, device="cuda") * 2 - 1) * 1.5
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
    C.set_option("dbg", dbg)
    r = []
    for prec in (3, 1):
        r.append(t(lambda: ops.udf_forward(net, prec, pts=x)))
        r.append(t(lambda: ops.udf_forward_grad(net, prec, pts=x[: P // 4])))
    print(f"dbg={dbg} (1=noMMA 2=noCopy 4=noEpiMath)  fwd3 {r[0]:.2f} ms  grad3(P/4) {r[1]:.2f} ms  fwd1 {r[2]:.2f} ms  grad1(P/4) {r[3]:.2f} ms", flush=True)
C.set_option("dbg", 0)
for cl in (1, 2, -2):
    C.set_option("cluster", cl)
    r = []
    for prec in (3, 1):
        r.append(t(lambda: ops.udf_forward(net, prec, pts=x)))
        r.append(t(lambda: ops.udf_forward_grad(net, prec, pts=x[: P // 4])))
    print(f"cluster={cl}  fwd3 {r[0]:.2f} ms  grad3(P/4) {r[1]:.2f} ms  fwd1 {r[2]:.2f} ms  grad1(P/4) {r[3]:.2f} ms", flush=True)
C.set_option("cluster", 1)
