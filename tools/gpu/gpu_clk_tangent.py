"""clock64 timeline of one tile of the tangent forward (mlp_kernel<1, MODE 3>: block 0, its 30th tile): epilogue warp 0
and the MMA-issuing warp, per layer.  usage: python tools/gpu/gpu_clk_tangent.py"""
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
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
C.set_option("dbg_iter", 30)


def run():
    C.check(L.emap_bwd_tangent_forward(desc, C.ptr(net.packed), C.ptr(x), None, None, None, 0, P, C.ptr(gbar),
                                       C.ptr(scales), C.ptr(stash[0]), C.ptr(stash[1]), st))


run()
torch.cuda.synchronize()
L.emap_debug_set_clk_buffer(C.ptr(buf))
run()
torch.cuda.synchronize()
L.emap_debug_set_clk_buffer(None)
C.set_option("dbg_iter", 1)
b = buf.cpu()
t0 = int(b[0])
print("tangent forward, block 0, tile iteration 30 (clk since the tile's first accumulator wait)")
print("  epilogue warp 0: [waiting, acc complete | chunk 0: tcgen05.ld done, math+stores done, handed off | chunk 1: same]")
print("  issuer: [layer start, acc free, K chunk 0..3 ready, all MMAs issued]")
for l in range(8):
    e = [int(v) - t0 if v else -1 for v in b[l * 8:l * 8 + 8]]
    m = [int(v) - t0 if v else -1 for v in b[72 + l * 8:72 + l * 8 + 7]]
    print(f"  L{l}: epi {e}   mma {m}", flush=True)
