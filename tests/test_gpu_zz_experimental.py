"""GPU, NON-GATING: runs the parity tests of the kernels that were written after round 1's GPU budget was
spent (K1r, shared-forward backward / tangent-only forward, K1 MODE 5, two-tile reverse sweep) -- each group in
its OWN child process, so that a trapped launch (a protocol bug traps after a bounded mbarrier wait) cannot
poison the CUDA context of the validated suite -- and reports the outcome without gating on it:

    child green  -> this test passes (and prints the child's summary)
    child red    -> pytest.xfail with the child's summary line: recorded, visible, not a suite failure.

These kernels are opt-in in the product; their tests proper are in tests/test_gpu_rgrad.py and
tests/test_gpu_rev2.py and run directly with EMAP_EXPERIMENTAL=1.  This file sorts last on purpose: every
validated GPU test has run before the first experimental launch.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GROUPS = [
    ("k1r", ["tests/test_gpu_rgrad.py", "-k", "not k1_dot and not shared"]),
    ("shared_backward", ["tests/test_gpu_rgrad.py", "-k", "shared"]),
    ("k1_dot", ["tests/test_gpu_rgrad.py", "-k", "k1_dot"]),
    ("rev2", ["tests/test_gpu_rev2.py"]),
]


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("name,args", GROUPS, ids=[g[0] for g in GROUPS])
def test_experimental_group_in_child_process(name, args, capsys):
    if os.environ.get("EMAP_EXPERIMENTAL") == "1":
        pytest.skip("the experimental tests are running directly in this session")
    import torch
    torch.cuda.empty_cache()                       # hand this process's cached blocks back before the child allocates
    env = dict(os.environ, EMAP_EXPERIMENTAL="1")
    cmd = [sys.executable, "-m", "pytest", "-q", "-rfEs", "-p", "no:cacheprovider"] + args
    try:
        res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
        out, rc = res.stdout + res.stderr, res.returncode
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
        rc = -1
    lines = [ln for ln in out.strip().splitlines() if ln.strip()]
    summary = lines[-1] if lines else "(no output)"
    tail = "\n".join(lines[-25:])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"experimental_{name}.log"), "w") as f:
        f.write(out)
    with capsys.disabled():
        print(f"\n[experimental:{name}] rc={rc}  {summary}")
        if rc != 0:
            print(tail)
    if rc != 0:
        pytest.xfail(f"experimental group '{name}' (opt-in kernels, not yet validated): {summary}")
