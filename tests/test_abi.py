"""CPU: the C-ABI library loads, exports every symbol include/fmrx.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fmrx
from util import assert_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fmrx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmrx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 35
    lib = C.CDLL(fmrx.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fmrx.h but not exported by libfmrx.so"
        assert n in fmrx.SIGNATURES, f"{n} has no ctypes signature in fmrx.SIGNATURES"
    assert set(fmrx.SIGNATURES) == set(names)
    assert fmrx.lib().fmrx_version() == 101


def test_struct_layouts_match_header():
    assert C.sizeof(fmrx.Config) == 32 and C.sizeof(fmrx.RdsEvent) == 16 and fmrx.EVENT_DTYPE.itemsize == 16
    assert C.sizeof(fmrx.Outputs) == 6 * C.sizeof(C.c_void_p)


def test_design_is_host_side_and_bit_identical(golden):
    """Filter design is host code in the product library; its taps must equal the reference's bit for bit."""
    g = golden["functions"]
    assert_bits(fmrx.design_lpf(2.4e6, 1e5, 151), g["lpf_rf0"], "rf lpf")
    assert_bits(fmrx.design_lpf(6e6, 16000, 3624), g["lpf_mono1"], "mode-1 lpf (NaN tap)")
    assert_bits(fmrx.design_lpf(float(np.float32(240000) * np.float32(19)), 28500, 2869), g["lpf_anti"], "anti-image")
    assert_bits(fmrx.design_lpf(240000 * 147, 16000, 151 * 147), g["lpf_441"], "44.1k")
    for name, args in (("bpf_pilot0", (18.5e3, 19.5e3, 240000)), ("bpf_stereo0", (22e3, 54e3, 240000)), ("bpf_pilot1", (18.5e3, 19.5e3, 6e6)),
                       ("bpf_stereo1", (22e3, 54e3, 6e6)), ("bpf_rds", (54000, 60000, 240000)), ("bpf_sq", (113500, 114500, 240000))):
        assert_bits(fmrx.design_bpf(*args, 151), g[name], name)
    assert_bits(fmrx.design_rrc(57000, 151), g["rrc"], "rrc")


def test_argument_errors_do_not_need_a_gpu():
    lib = fmrx.lib()
    assert lib.fmrx_design_lpf(1.0, 1.0, 0, None) == -1
    assert b"bad argument" in lib.fmrx_last_error()
    h = C.c_void_p()
    cfg = fmrx.Config(7, 0, 1, 1, 0, 0, 0, 0)
    assert lib.fmrx_batch_create(C.byref(cfg), C.byref(h)) == -1 and not h.value


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry fails loudly (FMRX_ERR_CUDA); nothing is computed on the host."""
    if fmrx.lib().fmrx_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fmrx.FmrxError, match="no CUDA device|CUDA|cuda"):
        fmrx.Batch(1)
    with pytest.raises(fmrx.FmrxError):
        fmrx.unpack_iq(np.arange(16, dtype=np.uint8))
    zi = np.zeros(150, np.float32)
    with pytest.raises(fmrx.FmrxError):
        fmrx.fir_decim(np.zeros(1000, np.float32), np.zeros(151, np.float32), zi, 5)


def test_format_block_matches_reference_lines(golden):
    ev = np.array([(3, 1, 1, 236), (3, 2, -1, 236), (3, 0, 2, 262)], fmrx.EVENT_DTYPE)
    assert fmrx.rds_format_block(3, 23, ev) == (" \n****************Prcoessing Block: 3****************\nFalse positive Syndrome B at position 236\n"
                                                 "~~~~~Re-Sync~~~~~\nSyndrome C at position 262\n")
    assert fmrx.rds_format_block(0, 23, ev[:0]).startswith("initial offset for clock recovery = 23\n \n")
