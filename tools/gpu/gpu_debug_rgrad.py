"""Bring-up diagnostic for K1r (mlp_rg.cu): per-MMA-step accumulator errors of tile 0 against the float64
emulation, for both precision modes, then output errors against K1g and the oracle on a ragged multi-tile
problem.  Prints everything it finds instead of stopping at the first mismatch.
usage: python tools/gpu/gpu_debug_rgrad.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C  # noqa: E402
from oracle import emap_oracle as O  # noqa: E402
from tests.helpers import oracle_params  # noqa: E402
from tests.test_rg_emulation import _emulate  # noqa: E402

STEP_NAMES = [f"fwd layer {l}" for l in range(8)] + [f"rev layer {l}" for l in range(7, -1, -1)]

for multires in (10, 6):
    p = oracle_params(True, multires)
    net = ops.PackedNet(multires)
    net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).cuda())
    torch.manual_seed(3)
    x = (torch.rand(128, 3) * 2 - 1) * 0.9
    steps = []
    eu, eg = _emulate(p, x, steps=steps)
    for prec, name in ((C.PREC_FP32X3, "fp32x3"), (C.PREC_HALF, "fp16")):
        try:
            udf, grad, dbg = ops.debug_rgrad(net, prec, x.cuda())
            torch.cuda.synchronize()
        except Exception as e:                                   # a trapped launch poisons the context: stop here
            print(f"[multires {multires} {name}] launch failed: {e}", flush=True)
            sys.exit(1)
        got = dbg.cpu().double().numpy()
        print(f"--- multires {multires}, {name}: max |acc - emulation| per MMA step (tile 0)")
        for i, want in enumerate(steps):
            err = np.abs(got[i] - want)
            r, c = np.unravel_index(np.argmax(err), err.shape)
            print(f"  step {i:2d} ({STEP_NAMES[i]:12s}): err {err.max():.3e}  scale {np.abs(want).max():.3e}  "
                  f"worst at row {r} col {c}  nan {int(np.isnan(got[i]).sum())}", flush=True)
        print(f"  udf  err vs emulation {np.abs(udf.cpu().numpy() - eu).max():.3e}")
        print(f"  grad err vs emulation {np.abs(grad.cpu().numpy() - eg).max():.3e}", flush=True)

p = oracle_params(True)
net = ops.PackedNet(10)
net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).cuda())
B, n = 6011, 11
o, d = O.synthetic_rays(B)
z = torch.rand(B, n) * 3 + 0.5
args = dict(rays_o=o.cuda(), rays_d=d.cuda(), z=z.cuda())
uf, gf = ops.udf_forward_grad(net, C.PREC_FP32X3, mode="forward", **args)
for flags in (0, 1, 2):
    C.set_option("rg_flags", flags)
    try:
        ur, gr = ops.udf_forward_grad(net, C.PREC_FP32X3, mode="reverse", **args)
        torch.cuda.synchronize()
        print(f"66,121 ragged points, rg_flags={flags}: |udf - K1g| {float((ur - uf).abs().max()):.3e}  "
              f"|grad - K1g| {float((gr - gf).abs().max()):.3e}", flush=True)
    except Exception as e:
        print(f"rg_flags={flags}: launch failed: {e}", flush=True)
        sys.exit(1)
    finally:
        C.set_option("rg_flags", 28)
