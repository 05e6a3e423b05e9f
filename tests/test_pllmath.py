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


@pytest.mark.parametrize("variant", [0, 1, 2], ids=["conversion instructions", "integer-built conversions", "same, theta0 from the staged table"])
def test_loop_bit_identical_to_libm_loop(checker, variant):
    """8 loops (19 kHz x2 and 114 kHz x0.5; clean, weak, detuned+strong, noisy with exact zeros) x 12 blocks: every
    carried float (integrator, phase estimate, feedback I/Q) and every NCO sample equal to the libm-only recurrence.
    Variant 0 is pll_step_fast, variants 1 / 2 pll_step_fast1 (what the kernel runs by default), the latter with the theta0
    table staged as the kernel stages it in shared memory."""
    r = json.loads(subprocess.check_output([checker, "loop", "12", str(variant)]))
    assert r["fast_steps"] > 0.99 * r["steps"], r   # the short-chain path is the one being exercised
    assert r["mismatches"] == 0, r


def test_integer_built_widening_is_the_cast(checker):
    """float -> double by re-packing the bits (pllmath::widen / widen_pos) equals the conversion for every normal finite float tried
    (4 M random bit patterns inside the range `ok` admits, plus the edges of that range)."""
    r = json.loads(subprocess.check_output([checker, "widen", "4000000"]))
    assert r["bad"] == 0, r
