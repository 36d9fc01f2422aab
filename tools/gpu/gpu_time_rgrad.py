"""K1r (reverse-mode gradient kernel) vs K1g (forward-mode): agreement + timing on 1 M points.
usage: python tools/gpu/gpu_time_rgrad.py"""
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


for prec, name in ((3, "fp32x3"), (1, "fp16")):
    # (independent of K1r: first, so that a K1r failure does not lose it)
    f0 = t(lambda: ops.udf_forward(net, prec, pts=x))
    uf, gf = ops.udf_forward_grad(net, prec, pts=x, mode="forward")
    ur, gr = ops.udf_forward_grad(net, prec, pts=x, mode="reverse")
    torch.cuda.synchronize()
    print(f"{name}: |udf_rev - udf_fwd| max {float((ur - uf).abs().max()):.3e}   "
          f"|grad_rev - grad_fwd| max {float((gr - gf).abs().max()):.3e}", flush=True)
    tf = t(lambda: ops.udf_forward_grad(net, prec, pts=x, mode="forward"))
    tr = t(lambda: ops.udf_forward_grad(net, prec, pts=x, mode="reverse"))
    C.set_option("rg_flags", 1)
    tr1 = t(lambda: ops.udf_forward_grad(net, prec, pts=x, mode="reverse"))
    C.set_option("rg_flags", 28)
    print(f"{name}: K1r with the N-split tail (rg_flags=1): {tr1:.2f} ms", flush=True)
    C.set_option("rg_flags", 2)
    tr2 = t(lambda: ops.udf_forward_grad(net, prec, pts=x, mode="reverse"))
    C.set_option("rg_flags", 28)
    print(f"{name}: K1r with the persisting-L2 window on the sigma scratch (rg_flags=2): {tr2:.2f} ms", flush=True)
    for fl in (0, 10, 12, 4):
        C.set_option("rg_flags", fl)
        ud, gd = ops.udf_forward_grad(net, prec, pts=x, mode="reverse")
        torch.cuda.synchronize()
        same = bool(torch.equal(ud, ur) and torch.equal(gd, gr))
        trd = t(lambda: ops.udf_forward_grad(net, prec, pts=x, mode="reverse"))
        C.set_option("rg_flags", 28)
        print(f"{name}: K1r rg_flags={fl} (default 8 = dynamic tiles; 0 = static round robin, 2 = persisting L2, 4 = rolled issuer loop): {trd:.2f} ms, bit-identical to the default: {same}", flush=True)
    flop = 2 * 918016 * P
    print(f"{name}: K1g {tf:.2f} ms ({flop / tf / 1e9:.0f} TF/s alg)   K1r {tr:.2f} ms ({flop / tr / 1e9:.0f} TF/s alg)   "
          f"K1 forward only {f0:.2f} ms", flush=True)
