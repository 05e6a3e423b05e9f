"""include/fmrx_dropin.hpp: the reference's own function names (src/filter.h, helper.h, rf_module.h) over the C-ABI.
CPU: the header compiles and links against libfmrx.so.  GPU: a program written like the reference's thread bodies
(tests/native/dropin_check.cpp: rf_thread, mono_stero_thread and rds_thread) reproduces the oracle's demod / mono / pilot / NCO /
stereo and the whole RDS branch (band, pllCombine's filter and NCO, mixer filter, 19/80 resampler, RRC) bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from fmrx import synth
from oracle import Chain
from util import assert_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "real-time-software-defined-radio_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin") / "dropin_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "native", "dropin_check.cpp"), "-o", out, "-L", PKG, "-lfmrx", f"-Wl,-rpath,{PKG}"])
    return out


def test_header_compiles_and_links(exe):
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_reference_style_program_matches_oracle(exe, tmp_path):
    nblk = 3
    raw = synth.synth_iq(nblk, 0, seed=11)
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.f32")
    raw.tofile(fin)
    subprocess.check_call([exe, fin, fout])
    got = np.fromfile(fout, np.float32).reshape(nblk, -1)
    sizes = [15360, 3072, 15360, 15360, 3072, 15360, 15360, 15360, 15360, 3648, 3648]
    chain = Chain(0, 1)  # intent profile: stereo computed in every block, outputs assigned
    for b in range(nblk):
        chain.block(raw[b * 307200:(b + 1) * 307200])
        off = 0
        for name, n in zip(("demod", "mono", "pilot", "nco", "stereo", "rds_bpf", "rds_sq", "rds_nco", "rds_lpf", "rds_res", "rds_rrc"), sizes):
            assert_bits(got[b, off:off + n], chain.tap(name)[:n], f"block {b} {name}")
            off += n
