"""A short run of the randomised parity soak (tools/fuzz_parity.py) inside the GPU suite: random sizes, depths, layouts, alignments and
representations against the oracle, a fixed seed so a failure reproduces."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_parity_short_run(seed):
	r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "8", str(seed)], capture_output=True, text=True, timeout=300)
	assert r.returncode == 0 and "no mismatch" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
