"""bf16 network: udf / gradient error of K1g (forward-mode) and K1r (reverse-mode) against the reference fixture,
next to the fp16 single-MMA and fp32x3 numbers.  usage: python tools/gpu/gpu_diag_bf16.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C  # noqa: E402
from tests.conftest import load_golden  # noqa: E402
from tests.helpers import oracle_params  # noqa: E402

g = load_golden("mlp_pert")
p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
x = g["x"].cuda()
for elem in ("fp16", "bf16"):
    net = ops.PackedNet(10, elem_type=elem)
    net.fold(flat)
    for prec, pname in ((C.PREC_FP32X3, "x3"), (C.PREC_HALF, "x1")):
        for mode in ("forward", "reverse"):
            u, gr = ops.udf_forward_grad(net, prec, pts=x, mode=mode)
            torch.cuda.synchronize()
            du = (u.cpu() - g["out"].reshape(-1)).abs()
            dg = (gr.cpu() - g["grad"].reshape(-1, 3)).abs().max(dim=1)[0]
            print(f"{elem} {pname} {mode:8s}: udf err max {float(du.max()):.2e}  grad err median {float(dg.median()):.2e} "
                  f"p90 {float(torch.quantile(dg, 0.9)):.2e} max {float(dg.max()):.2e}", flush=True)
