"""GPU: every kernel family in every configuration (modes x numerics x RDS back ends x quality profile, ragged batch, checkpoint /
resume, ring, function-level operators at sizes with partial tiles) -- tools/memcheck_chain.py, which is also the driver for
compute-sanitizer where that is available.  Across configurations the int16 audio of REFERENCE and STRICT numerics and of both RDS
back ends must be identical and the RDS bits of every configuration must agree."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_every_configuration_runs_and_agrees():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "memcheck_chain.py")], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "all ok" in r.stdout
