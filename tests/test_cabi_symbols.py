"""CPU: the C-ABI library loads and exports every symbol include/emap_b200.h declares; host-only
entry points work without a GPU; the product refuses to run without a device."""
import ctypes
import os
import re

import pytest
import torch

from emap_b200 import _cabi as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "emap_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(emap_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 18
    raw = ctypes.CDLL(C.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in emap_b200.h but not exported"
        assert n in C.SIGNATURES, f"{n} has no ctypes signature in _cabi.py"
    for n in C.SIGNATURES:
        assert n in names, f"{n} bound in _cabi.py but not declared in the header"


def test_host_only_entry_points():
    L = C.lib()
    assert L.emap_abi_version() == 2
    for mr, expect in ((10, 462980), (6, 462980 - 256 * 24 + 0), (0, None)):
        d = C.NetDesc(mr, 0, 1.0, 0)
        n = L.emap_flat_param_count(ctypes.byref(d))
        pe = 3 + 6 * mr
        ref = 256 * (pe + 2) + 2 * 256 * 258 + (256 - pe) * 258 + 4 * 256 * 258 + 258
        assert n == ref
        assert L.emap_packed_size(ctypes.byref(d)) > 128 * 1024 * 16
    bad = C.NetDesc(11, 0, 1.0, 0)
    assert L.emap_flat_param_count(ctypes.byref(bad)) == 0
    assert b"multires" in L.emap_last_error()
    assert L.emap_set_option(b"cluster", 4) != 0
    assert L.emap_set_option(b"cluster", 1) == 0


def test_no_cpu_fallback():
    from emap_b200.udf_model import UDFNetwork
    torch.manual_seed(0)
    net = UDFNetwork(3, 1, 256, 8, skip_in=[4], multires=10)
    with pytest.raises(RuntimeError):
        net(torch.zeros(4, 3))                       # CPU tensors: refuse, do not emulate
    with pytest.raises(NotImplementedError):
        UDFNetwork(3, 1, 128, 8, skip_in=[4], multires=10)
    # SURVEY 8f modules: same rule
    from emap_b200.extract_pointcloud import get_udf_normals_grid, get_udf_normals_slow
    from emap_b200.ray_sampler import RaySampler
    with pytest.raises(RuntimeError):
        get_udf_normals_grid(net.udf, net.gradient, 4, 0.1, device="cpu")
    with pytest.raises(RuntimeError):
        get_udf_normals_slow(net.udf, net.gradient, 0.1, torch.zeros(2, 3), False, device="cpu")
    with pytest.raises(RuntimeError):
        RaySampler(torch.zeros(1, 4, 4, 1), torch.eye(4)[None], torch.eye(4)[None], device="cpu")
    from emap_b200 import ops
    with pytest.raises(RuntimeError):
        ops.null_direction(torch.zeros(2, 5, 3))
    # the CUDA-graph wrapper and the standalone RenderingNetwork operator: same rule
    from emap_b200.graph import GraphedStep
    from emap_b200.udf_model import RenderingNetwork
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            GraphedStep(lambda x: (x,), [torch.zeros(1)])
    rn = RenderingNetwork(d_feature=8, mode="no_normal", d_in=6, d_out=1, d_hidden=16, n_layers=1)
    with pytest.raises(RuntimeError):
        rn(torch.zeros(2, 3), None, torch.zeros(2, 3), torch.zeros(2, 8))


def test_state_dict_keys_and_init_match_reference(golden):
    from emap_b200.udf_model import BetaNetwork, SingleVarianceNetwork, UDFNetwork
    torch.manual_seed(0)
    net = UDFNetwork(d_in=3, d_out=1, d_hidden=256, n_layers=8, skip_in=[4], multires=10, bias=0.5,
                     scale=1.0, geometric_init=True, weight_norm=True, udf_type="abs")
    ref = golden("net_init_state")
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(sd[k], ref[k]), k          # same seed -> bit-identical initial weights
    assert [tuple(p.shape) for p in net.parameters()][:3] == [(256,), (256, 1), (256, 63)]
    assert sorted(SingleVarianceNetwork(0.3).state_dict()) == ["second_variance", "variance"]
    assert sorted(BetaNetwork().state_dict()) == ["beta", "gamma", "zeta"]
    g = golden("scalars")
    assert torch.equal(SingleVarianceNetwork(0.3)(torch.zeros(1, 3)), g["inv_s"])
    b = BetaNetwork(0.5, 0.3, 0.3, 5e-5, True, True, False)
    assert torch.equal(b.get_beta(), g["beta"]) and torch.equal(b.get_gamma(), g["gamma"])
