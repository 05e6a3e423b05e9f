"""GPU: the fm_radio executable's process contract against the reference executable's recorded stdout / stderr."""
import subprocess

import numpy as np
import pytest

import fmrx
from fmrx import synth
from util import assert_bits

pytestmark = pytest.mark.gpu


def run_cli(args, raw):
    r = subprocess.run([fmrx.CLI_PATH] + args, input=raw.tobytes(), capture_output=True, timeout=600)
    return np.frombuffer(r.stdout, np.int16), r.stderr.decode(), r.returncode


@pytest.mark.parametrize("mode,extra", [(0, []), (0, ["--blocks", "3"]), (1, [])])
def test_cli_matches_reference_binary(golden, mode, extra):
    g = golden[f"chain_mode{mode}"]
    nblk = int(g["nblk"])
    raw = synth.synth_iq(nblk, mode, seed=int(g["seed"]))
    tail = np.zeros(1000, np.uint8)  # a trailing partial block is ignored (Q9, normalised)
    audio, err, rc = run_cli((["1"] if mode == 1 else []) + extra, np.concatenate([raw, tail]))
    assert rc == 0, err
    assert_bits(audio, g["binary_audio"], "stdout vs reference fm_radio")
    lines = err.splitlines()
    assert lines[0] == str(1 + (mode == 1)) and lines[1] == f"Operating in mode {mode}" and lines[2] == f"rf_Fs = {2500000 if mode else 2400000}"
    assert lines[-1].startswith("Run: gnuplot")
    if mode == 0:
        body = "\n".join(lines[3:-1]) + "\n"
        assert body == str(g["binary_frame_text"])


@pytest.mark.parametrize("extra", [[], ["--numerics", "reference"]])
def test_cli_matches_reference_on_noisy_input(extra):
    """The same contract on an input without wide decision margins: AWGN at 4 dB CNR.  Golden (tests/golden/make_noisy_golden.py):
    stdout of the unmodified reference executable, and the frame_thread lines the reference's own functions give when driven in
    sequence -- the executable's stderr is not deterministic on such input (four runs, four texts), its audio is.  Default
    settings of this executable (STRICT numerics, stage-by-stage RDS back end: every rounding is the reference's): both streams byte
    for byte.  With `--numerics reference` (the batch library's default: FFMA in pllCombine's filter, symbol-rate back end) the audio
    is still byte-identical and the RDS lines are compared too -- on this input they are equal; in general a near-tie may flip (2 of
    1.2 M bits)."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_mode0_noisy.npz"), allow_pickle=False)
    nblk = int(g["nblk"])
    raw = synth.synth_iq(nblk, 0, seed=int(g["seed"]), cnr_db=float(g["cnr_db"]), noise_seed=int(g["noise_seed"]))
    from util import sha

    assert sha(raw) == str(g["input_sha256"])
    audio, err, rc = run_cli(extra, raw)
    assert rc == 0, err
    assert_bits(audio, g["binary_audio"], "stdout vs reference fm_radio on noisy input")
    body = "\n".join(err.splitlines()[3:-1]) + "\n"
    assert body == str(g["frame_text"]), "frame_thread lines vs the reference's functions on noisy input"


def test_cli_rejects_bad_modes_like_the_reference():
    for arg, msg in (("0", "Wrong mode 0"), ("2", "Wrong mode 2"), ("x", "Wrong mode 0")):
        _, err, rc = run_cli([arg], np.zeros(0, np.uint8))
        assert rc == 1 and msg in err


def test_cli_intent_profile(golden):
    g = golden["chain_mode0"]
    raw = synth.synth_iq(int(g["nblk"]), 0, seed=int(g["seed"]))
    audio, err, rc = run_cli(["--profile", "intent", "--quiet"], raw)
    assert rc == 0 and "Syndrome" not in err
    assert_bits(audio, g["intent_audio"], "intent profile stdout")


def test_cli_rds_info_reports_the_programme():
    """extension beyond the reference CLI: --rds-info runs the RDS application layer on the decoded bits"""
    raw = synth.synth_iq(40, 0, seed=1, rds_payload=lambda n: synth.rds_group_bits(0xBEEF, "CLI TEST", "radiotext through the executable", pty=3, n_bits=n))
    _, err, rc = run_cli(["--profile", "intent", "--quiet", "--rds-info", "--blocks", "4"], raw)
    assert rc == 0, err
    line = [l for l in err.splitlines() if l.startswith("RDS: ")]
    assert len(line) == 1 and 'PI BEEF PTY 3 TP 0 PS "CLI TEST" RT "radiotext through the executable"' in line[0], err


def test_cli_empty_and_short_input():
    """Q9, normalised: only whole 307200-byte blocks are processed -- no input or less than one block gives no audio,
    exit 0, and still the reference's banner lines."""
    for n in (0, 1000, 307199):
        audio, err, rc = run_cli([], np.zeros(n, np.uint8))
        assert rc == 0 and audio.size == 0, (n, err)
        lines = err.splitlines()
        assert lines[0] == "1" and lines[1] == "Operating in mode 0" and lines[2] == "rf_Fs = 2400000"


def test_cli_rejects_what_the_reference_rejects():
    """src/fm_radio.cpp:736-764: exactly "1" selects mode 1; "0", "2", words and unknown options end with exit code 1"""
    for args in (["0"], ["2"], ["stereo"], ["--no-such-option"], ["1", "--audio-rate", "12345"]):
        r = subprocess.run([fmrx.CLI_PATH] + args, input=b"", capture_output=True, timeout=60)
        assert r.returncode == 1 and not r.stdout, args
