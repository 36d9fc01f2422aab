import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C
from tests.helpers import oracle_params
p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
net = ops.PackedNet(10); net.fold(flat)
buf = torch.zeros(176, dtype=torch.int64, device="cuda")
C.lib().emap_debug_set_clk_buffer(C.ptr(buf))
for prec in (3, 1):
  for mode, P in ((0, 148 * 128 * 3), (1, 148 * 32 * 3)):
    x = (torch.rand(P, 3, device="cuda") * 2 - 1) * 1.5
    buf.zero_()
    ops.debug_mlp(net, prec, mode, x); torch.cuda.synchronize()
    full = buf.cpu()
    b = full[:144].reshape(2, 9, 8)
    t0 = int(b[0, 0, 0])
    print(f"=== prec={prec} mode={mode}: epilogue warp0 [wait_start, acc_full, ld0, math0, arrive0, ld1, math1, arrive1]  |  MMA role0 [start, acc_empty, a0, a1, a2, a3, committed]")
    for l in range(9):
        e = [int(v) - t0 if v else -1 for v in b[0, l]]
        m = [int(v) - t0 if v else -1 for v in b[1, l, :7]]
        print(f"  L{l}: epi {e}   mma {m}")
    ex = [int(v) - t0 if v else -1 for v in full[144:164]]
    print("  MMA role0 L2 kc1 items [wait_start, full_seen, issued, committed] x parts:", ex[:8])
    print("  producer items 24..27 [start, empty_seen, copy_issued]:", ex[8:20])
