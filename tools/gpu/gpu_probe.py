"""One experiment per process (a device trap kills the CUDA context).  usage: gpu_probe.py MODE PREC P [cluster]"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from emap_b200 import ops, _cabi as C
from oracle import emap_oracle as O
from tests.helpers import oracle_params
mode, prec, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cl = int(sys.argv[4]) if len(sys.argv) > 4 else 1
C.set_option("cluster", cl)
p = oracle_params(True)
flat = torch.cat([t.reshape(-1) for t in p.tensors()]).cuda()
net = ops.PackedNet(10); net.fold(flat)
g = torch.Generator().manual_seed(5)
x = (torch.rand(P, 3, generator=g) * 2 - 1) * 1.5
ref = O.udf_forward(p, x)[0][:, 0]
if mode == 0:
    u, _ = ops.udf_forward(net, prec, pts=x.cuda())
    torch.cuda.synchronize()
    print(f"OK mode=0 prec={prec} P={P} cl={cl} udf_err={(u.cpu()-ref).abs().max().item():.3e}")
else:
    u, gr = ops.udf_forward_grad(net, prec, pts=x.cuda())
    torch.cuda.synchronize()
    rg = O.udf_gradient(p, x).detach()
    print(f"OK mode=1 prec={prec} P={P} cl={cl} udf_err={(u.cpu()-ref).abs().max().item():.3e} grad_err={(gr.cpu()-rg).abs().max().item():.3e} gmax={rg.abs().max().item():.2f}")
