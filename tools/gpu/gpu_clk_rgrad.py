"""clock64 timeline of one K1r tile (block 0, its second tile): epilogue warp 0 and the MMA-issuing warp, per
MMA step.  usage: python tools/gpu/gpu_clk_rgrad.py"""
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
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
C.lib().emap_debug_set_clk_buffer(C.ptr(buf))
names = [f"fwd L{l}" for l in range(8)] + [f"rev L{l}" for l in range(7, -1, -1)]
# a steady-state tile: full-size launch (1 M points, 56 tiles per CTA), the 30th tile of block 0
C.set_option("dbg_iter", 30)
for flags in (8,):
    C.set_option("rg_flags", flags)
    for prec in (3, 1):
        x = (torch.rand(1 << 20, 3, device="cuda") * 2 - 1) * 1.5
        buf.zero_()
        ops.debug_rgrad(net, prec, x)
        torch.cuda.synchronize()
        b = buf.cpu()
        t0 = int(b[0])
        print(f"=== rg_flags={flags} prec={prec}: epilogue warp 0 [waiting, acc complete, chunk 0 handed off, done]  |  "
              f"issuer [step start, acc free, first K chunk ready, all MMAs issued]   (clk since the tile's first wait)")
        for s in range(15):
            e = [int(v) - t0 if v else -1 for v in b[4 * s:4 * s + 4]]
            m = [int(v) - t0 if v else -1 for v in b[64 + 4 * s:64 + 4 * s + 4]]
            print(f"  step {s:2d} {names[s]}: epi {e}   mma {m}", flush=True)
        m = [int(v) - t0 if v else -1 for v in b[64 + 60:64 + 64]]
        print(f"  step 15 {names[15]}: mma {m}", flush=True)
        per = b[256:256 + 2 * 148].reshape(148, 2)
        cyc = per[:, 0].double()
        order = torch.argsort(cyc)
        print(f"  per-block elapsed cycles: min {cyc.min():.0f} median {cyc.median():.0f} max {cyc.max():.0f} "
              f"(max/min {cyc.max() / cyc.min():.3f}); slowest blocks (block, smid, Mclk): "
              f"{[(int(i), int(per[i, 1]), round(float(cyc[i]) / 1e6, 2)) for i in order[-6:]]}; fastest: "
              f"{[(int(i), int(per[i, 1]), round(float(cyc[i]) / 1e6, 2)) for i in order[:6]]}", flush=True)
C.set_option("rg_flags", 28)
C.set_option("dbg_iter", 1)
C.lib().emap_debug_set_clk_buffer(None)
