"""GPU: the two-tiles-in-flight reverse sweep (emap_b200/csrc/mlp_rev2.cu, emap_set_option("rev_tiles", 2))
must reproduce the validated kernel (mlp_rev.cu) BIT FOR BIT: same arithmetic per tile, only the
interleaving of two tiles per CTA differs.  Written after round 1's GPU budget was spent (protocol modelled
in tests/test_rev2_protocol.py); opt-in, runs with EMAP_EXPERIMENTAL=1 like tests/test_gpu_rgrad.py.
"""
import ctypes
import os

import pytest
import torch

from tests.helpers import oracle_params

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("EMAP_EXPERIMENTAL") != "1",
                       reason="rev2 not yet validated on hardware: set EMAP_EXPERIMENTAL=1 to run"),
    pytest.mark.timeout(120),
]


@pytest.mark.parametrize("P", [64, 1000, 148 * 128 + 77, 60011])
def test_rev2_bit_identical_to_rev(P):
    """P = one tile (Y empty), ragged, one pair per CTA + remainder, several pairs per CTA."""
    from emap_b200 import ops, _cabi as C
    p = oracle_params(True)
    net = ops.PackedNet(10)
    net.fold(torch.cat([t.reshape(-1) for t in p.tensors()]).cuda())
    torch.manual_seed(P)
    x = ((torch.rand(P, 3) * 2 - 1) * 0.9).cuda()
    gbar = (torch.randn(P, 3) * 0.1).cuda()
    L, desc, st = C.lib(), ctypes.byref(net.desc), C.stream()
    st_u0, st_u = ops.alloc_backward_stash(P, x.device)
    C.check(L.emap_bwd_dual_forward(desc, C.ptr(net.packed), C.PREC_HALF, C.ptr(x), None, None, None, 0, P,
                                    C.ptr(gbar), C.ptr(st_u0), C.ptr(st_u), st))
    coef = (torch.randn(2 * P) * 0.5).cuda()
    outs = []
    for tiles in (1, 2):
        st_a = torch.full((8, 2 * P, 256), float("nan"), dtype=torch.float16, device=x.device)
        try:
            C.set_option("rev_tiles", tiles)
            C.check(L.emap_bwd_reverse_sweep(desc, C.ptr(net.packed), C.ptr(coef), C.ptr(st_u), C.ptr(st_a), P, st))
            torch.cuda.synchronize()
        finally:
            C.set_option("rev_tiles", 1)
        outs.append(st_a)
    assert torch.isfinite(outs[1]).all()
    assert torch.equal(outs[0], outs[1])
