"""CPU: the bench harness measures what BASELINE.json names -- the workload presets against the config strings, one
workload description for both arms, and the clock sampler degrading gracefully where there is no GPU."""
import argparse
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_presets_match_baseline_configs():
    import bench
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    for key, idx in (("c2", 1), ("c3", 2), ("c4", 3), ("c5", 4)):
        text = cfgs[idx].replace("×", "x")
        w = bench.WORKLOADS[key]
        rays = int(re.search(r"(\d+) rays", text).group(1))
        assert w["rays"] == rays, (key, text)
        m = re.search(r"rays x \((\d+)\+(\d+) hierarchical\)", text)
        if m:
            assert (w["n0"], w["ni"]) == (int(m.group(1)), int(m.group(2))), (key, text)
        else:
            assert w["n0"] + w["ni"] == int(re.search(r"rays x (\d+) samples", text).group(1)), (key, text)
        if "bf16" in text:
            assert w.get("precision") == "bf16"
        if "fp32" in text:
            assert w.get("precision", "fp32") == "fp32"


def test_both_arms_describe_the_same_workload():
    import bench
    a = argparse.Namespace(rays=4096, scaling="weak", mode="train")
    b = argparse.Namespace(rays=4096, scaling="weak", mode="train")
    assert bench.workload_config(a, 1) == bench.workload_config(b, 1)
    s = bench.workload_config(argparse.Namespace(rays=4096, scaling="strong", mode="train"), 8)
    assert s["rays_per_gpu"] == 512 and s["samples_per_ray"] == 256


def test_clock_sampler_without_a_gpu_reports_unavailable():
    import torch
    import bench
    if torch.cuda.is_available():
        return
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert out["reasons"] == ["unavailable"] and out["sm_mhz"] is None
