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


_PLL_VARIANT_SNIPPET = r"""
import hashlib, os, sys
import numpy as np
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "real-time-software-defined-radio_b200"))
import fmrx
from fmrx import synth
S, B = 40, 3   # more than one warp of loops per side, a noisy and a clean half, three blocks in one call and one more after it
raw = np.stack([synth.synth_iq(B + 1, 0, seed=100 + s, cnr_db=(6.0 if s % 2 else None)) for s in range(S)])
h = hashlib.sha256()
with fmrx.Batch(S, mode=0, profile=fmrx.PROFILE_INTENT, max_blocks=B, device=0) as rx:
    for lo, hi in ((0, B), (B, B + 1)):
        res = rx.process(raw[:, lo * fmrx.BLOCK_BYTES:hi * fmrx.BLOCK_BYTES], want_float=True)
        for k in ("audio", "audio_f", "rds_bits", "rds_n_bits", "rds_events", "rds_n_events"):
            h.update(np.ascontiguousarray(res[k]).tobytes())
    h.update(np.ascontiguousarray(rx.get_state()).tobytes())
print("digest", h.hexdigest())
"""


def test_pll_step_variants_are_bit_identical(tmp_path):
    """FMRX_PLL_STEP selects the form of the PLL step (csrc/fmrx_pllmath.h: 0 conversion instructions, 1 / 2 integer-built widenings and
    theta0 for the next sample's sign; 2 is the default).  Same stations, clean and noisy, through the whole chain under each: float and
    int16 audio, RDS bits, sync events and the carried state (PLL floats included) must be the same bytes."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "variant.py"
    script.write_text(_PLL_VARIANT_SNIPPET)
    digests = {}
    for v in ("0", "1", "2"):
        r = subprocess.run([sys.executable, str(script), root], capture_output=True, text=True, timeout=600, env=dict(os.environ, FMRX_PLL_STEP=v))
        assert r.returncode == 0, (v, r.stdout[-500:], r.stderr[-2000:])
        digests[v] = [ln.split()[1] for ln in r.stdout.splitlines() if ln.startswith("digest")][0]
    assert digests["0"] == digests["1"] == digests["2"], digests
