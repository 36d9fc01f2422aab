"""Bring-up of mlp_dw.cu: the isolated contraction check (tests/test_gpu_dw.py) under a few candidate byte strides
of the MN-major operand descriptors, plus timing at 1 M points.  usage: python tools/gpu/gpu_probe_dw.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C  # noqa: E402
from tests.test_gpu_dw import JOBS, PART  # noqa: E402

dev = "cuda"
net = ops.PackedNet(10)
ws = ops._bwd_workspace(torch.device(dev))
sms = torch.cuda.get_device_properties(0).multi_processor_count


def check(P, lbo, sbo):
    C.set_option("dw_lbo", lbo)
    C.set_option("dw_sbo", sbo)
    torch.manual_seed(1)
    st_a = (torch.randn(8, 2 * P, 256, device=dev) * 0.5).half()
    st_u = (torch.randn(8, 2 * P, 256, device=dev) * 0.5).half()
    st_u0 = (torch.randn(2 * P, 64, device=dev) * 0.5).half()
    ws.zero_()
    n = int(C.lib().emap_bwd_weight_grads(ctypes.byref(net.desc), C.ptr(st_a), C.ptr(st_u0), C.ptr(st_u), P,
                                          C.ptr(ws), ws.numel(), C.stream()))
    torch.cuda.synchronize()
    wsf = ws.view(torch.float32)
    parts = wsf[:sms * PART].view(sms, PART)[:n].double().sum(0)
    dbp = wsf[sms * PART:sms * PART + sms * 2048].view(sms, 8, 256)[:n].double().sum(0)
    errs = []
    for off, nn, la, lu in JOBS:
        ref = st_a[la].double().t() @ (st_u0 if lu == "u0" else st_u[lu]).double()
        got = parts[off:off + 256 * nn].view(256, nn)
        errs.append(float((got - ref).abs().max()) / float(ref.abs().max()))
    dberr = max(float((dbp[l] - st_a[l][:P].double().sum(0)).abs().max()) for l in range(8))
    return n, errs, dberr


for lbo, sbo in ((8192, 1024), (1024, 8192), (8192, 128), (128, 1024)):
    try:
        n, errs, dberr = check(1000, lbo, sbo)
        print(f"lbo={lbo} sbo={sbo}: parts={n} rel errors per job {[f'{e:.1e}' for e in errs]}  db abs err {dberr:.1e}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"lbo={lbo} sbo={sbo}: FAILED {e!r}", flush=True)
        break
C.set_option("dw_lbo", 8192)
C.set_option("dw_sbo", 1024)

P = 1 << 20
st_a = torch.randn(8, 2 * P, 256, device=dev, dtype=torch.float16)
st_u = torch.randn(8, 2 * P, 256, device=dev, dtype=torch.float16)
st_u0 = torch.randn(2 * P, 64, device=dev, dtype=torch.float16)


def run():
    C.lib().emap_bwd_weight_grads(ctypes.byref(net.desc), C.ptr(st_a), C.ptr(st_u0), C.ptr(st_u), P, C.ptr(ws),
                                  ws.numel(), C.stream())


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
gb = (8 * 2 + 1 + 7) * 2 * P * 256 * 2 / 1e9 + 2 * 2 * P * 64 * 2 / 1e9
print(f"weight_grad_kernel at P = 1 M: {ms:.2f} ms for {gb:.1f} GB of stash = {gb / ms:.2f} TB/s", flush=True)
