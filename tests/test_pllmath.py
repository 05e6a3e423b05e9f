"""CPU check of the PLL's short-chain double-precision kernels (csrc/fmrx_pllmath.h, compiled for the host) against
glibc — the libm the reference's fmPLL links (/root/reference/src/helper.cpp:32-45 calls atan2 / cos / sin in double on
float-valued arguments and rounds each result to float).  The CUDA kernel compiles the very same header."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "pllmath_check.cpp")
INC = os.path.join(ROOT, "real-time-software-defined-radio_b200", "csrc")


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("pllmath") / "pllmath_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-I", INC, SRC, "-o", exe, "-lm"])
    return exe


def test_sincos_within_two_ulp_and_float_identical(checker):
    r = json.loads(subprocess.check_output([checker, "sincos", "4000000"]))
    # arguments are float-valued, 1e-3 .. 1e9 in magnitude, both signs
    assert r["max_ulp_sin"] <= 2.5 and r["max_ulp_cos"] <= 2.5, r
    assert r["float_flips"] == 0, r


def test_loop_bit_identical_to_libm_loop(checker):
    """8 loops (19 kHz x2 and 114 kHz x0.5; clean, weak, detuned+strong, noisy with exact zeros) x 12 blocks: every
    carried float (integrator, phase estimate, feedback I/Q) and every NCO sample equal to the libm-only recurrence."""
    r = json.loads(subprocess.check_output([checker, "loop", "12"]))
    assert r["fast_steps"] > 0.99 * r["steps"], r   # the short-chain path is the one being exercised
    assert r["mismatches"] == 0, r
