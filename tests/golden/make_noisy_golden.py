"""Generates tests/golden/chain_mode0_noisy.npz: what the UNMODIFIED reference computes on a NOISY synthetic multiplex -- white
Gaussian noise ahead of the 8-bit quantiser at 4 dB carrier-to-noise ratio over the RF rate, where the reference's own RDS bit error
rate is about 1 % above its floor, i.e. symbol decisions without wide margins.  The drop-in executable has to reproduce both streams
byte for byte (tests/test_gpu_cli.py::test_cli_matches_reference_on_noisy_input).

  * stdout (int16 audio): from the reference EXECUTABLE (oracle/_ref/fm_radio), which is deterministic in its audio;
  * the frame_thread lines of stderr: from the reference's own FUNCTIONS driven in sequence (oracle/_ref/libfmref.so: the unmodified
    objects behind oracle/ref_shim.cpp: rds_thread's calls, then frame_thread's body).  The executable itself cannot serve here:
    its RDS output is NOT DETERMINISTIC -- four runs on this input print four different sequences of syndromes (and on a CLEAN
    12-block input two of three runs differ), consistent with the hazards SURVEY App. A lists as Q16 (ring slot written before the
    lock is taken, `if`-guarded condition waits), observed here for the first time.  Audio is unaffected (identical md5 over every run).  The
    script prints how many distinct stderr texts a few runs of the executable give, for the record.

    make -C oracle && python tests/golden/make_noisy_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
sys.path.insert(0, HERE)

from fmrx import synth  # noqa: E402
from make_golden import frame_text, sha  # noqa: E402
from oracle import Ref, RefChain, run_ref_binary  # noqa: E402

NBLK, SEED, CNR_DB, NOISE_SEED = 12, 1, 4.0, 41


def main():
    raw = synth.synth_iq(NBLK, 0, seed=SEED, cnr_db=CNR_DB, noise_seed=NOISE_SEED)
    audio, err, rc = run_ref_binary(raw, 0)
    assert rc == 0
    texts = {frame_text(err, NBLK)}
    for _ in range(3):
        a2, e2, _ = run_ref_binary(raw, 0)
        assert np.array_equal(a2[:NBLK * 2 * 3072], audio[:NBLK * 2 * 3072]), "the executable's audio is deterministic"
        texts.add(frame_text(e2, NBLK))
    print(f"the reference executable printed {len(texts)} different frame_thread texts in 4 runs on this input")
    ch, rrcs, shim_audio = RefChain(0, 0), [], []
    for b in range(NBLK):
        shim_audio.append(ch.block(raw[b * 307200:(b + 1) * 307200]))
        rrcs.append(ch.taps["rds_rrc"])
    assert np.array_equal(np.concatenate(shim_audio), audio[:NBLK * 2 * 3072]), "reference functions in sequence != executable stdout"
    text = Ref().frame_thread(np.stack(rrcs))
    d = dict(nblk=NBLK, seed=SEED, cnr_db=CNR_DB, noise_seed=NOISE_SEED, input_sha256=sha(raw), binary_audio=audio[:NBLK * 2 * 3072], frame_text=text,
             executable_distinct_texts_in_4_runs=len(texts))
    np.savez_compressed(os.path.join(HERE, "chain_mode0_noisy.npz"), **d)
    print(d["frame_text"])
    print(os.path.getsize(os.path.join(HERE, "chain_mode0_noisy.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
