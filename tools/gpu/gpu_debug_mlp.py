"""GPU diagnostic (not a pytest): layer-by-layer comparison of the tcgen05 MLP kernel against the
fp64 oracle.  Run on the B200 box:  python tools/gpu/gpu_debug_mlp.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from emap_b200 import ops, _cabi as C  # noqa: E402
from oracle import emap_oracle as O  # noqa: E402
from tests.helpers import oracle_params  # noqa: E402
from tests.conftest import load_golden  # noqa: E402


def layer_preacts(p, x):
    """fp64 pre-activations of every layer, value rows only: list of [P,out]."""
    p = p.to(torch.float64)
    x = x.double()
    e = O.posenc(x, p.multires)
    W = O.effective_weights(p)
    h, pre = e, []
    for l in range(9):
        if l in p.skip_in:
            h = torch.cat([h, e], 1) / math.sqrt(2)
        a = torch.nn.functional.linear(h, W[l], p.b[l])
        pre.append(a)
        if l < 8:
            h = torch.nn.functional.softplus(a, beta=100)
    return pre


def main():
    torch.manual_seed(0)
    dev = "cuda"
    for pert in (False, True):
        p = oracle_params(pert)
        flat = torch.cat([t.reshape(-1) for t in p.tensors()]).to(dev)
        net = ops.PackedNet(10)
        net.fold(flat)
        torch.cuda.synchronize()
        g = load_golden("mlp_pert" if pert else "mlp_init")
        x = g["x"]
        pre = layer_preacts(p, x)
        for prec, name in ((C.PREC_FP32X3, "fp32x3"), (C.PREC_HALF, "fp16")):
            for mode in (0, 1):
                udf, grad, dbg = ops.debug_mlp(net, prec, mode, x.to(dev))
                torch.cuda.synchronize()
                dbg = dbg.cpu().double()
                print(f"--- pert={pert} prec={name} mode={mode}")
                for l in range(9):
                    od = pre[l].shape[1]
                    if mode == 0:
                        ref = pre[l][:128] - p.b[l].double()[None, :]
                        got = dbg[l, :, :od]
                    else:
                        # value rows of tile 0: row = 32q + 4p  <->  point q*8+p
                        rows = torch.tensor([32 * q + 4 * pp for q in range(4) for pp in range(8)])
                        ref = pre[l][:32] - p.b[l].double()[None, :]
                        got = dbg[l, rows, :od]
                    if l == 8:
                        got = got[:, :1]
                    err = (got - ref).abs().max().item()
                    print(f"  layer {l}: max|acc-ref| = {err:.3e}   (ref scale {ref.abs().max().item():.3f})")
                eu = (udf.cpu() - g["udf"][:, 0]).abs().max().item()
                print(f"  udf  max abs err vs reference: {eu:.3e}")
                if mode == 1:
                    eg = (grad.cpu() - g["grad"][:, 0]).abs().max().item()
                    print(f"  grad max abs err vs reference: {eg:.3e}  (|grad| max {g['grad'].abs().max().item():.3f})")


if __name__ == "__main__":
    main()
