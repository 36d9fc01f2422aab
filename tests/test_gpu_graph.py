"""GPU: one whole training iteration captured in a CUDA graph (emap_b200.graph.GraphedStep) reproduces the eager
iterations -- same losses, same parameters after three optimizer steps -- and an inference render() captured the
same way returns the eager result.  The path is deterministic (fixed-order partial sums, tile schedule does not
change results), so the comparison is tight: 1e-6 relative on the losses, 1e-6 absolute on the parameters.

The bodies live in tests/graph_cases.py and run in a child process each: a fault inside a graph replay is sticky
for the process and must not poison the rest of the suite."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case", ["train", "infer"])
def test_graphed_step(case):
    p = subprocess.run([sys.executable, "-m", "tests.graph_cases", case], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), (p.stdout[-2000:], p.stderr[-4000:])
