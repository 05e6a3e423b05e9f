"""Generates tests/golden/model_rds.npz by running the UNMODIFIED reference script model/fmRDSblock.py here (runpy, a stub
matplotlib, its hard-coded input path ../data/samples_rds_1029.raw provided in a scratch directory) on a seeded synthetic
multiplex, and captures what it computes:
  * its stdout (clock offset, Manchester screening counts, start position, every syndrome line) -- the text a drop-in has to
    reproduce;
  * per block, the arrays its loop body leaves in the module globals (rrc_rds, symbols_I, diff_bits, ...), snapshotted from
    inside the script's own `print('Processing block ...')` call at the top of the next iteration (the script is not edited:
    `print` is replaced through runpy's init_globals), and the final state after the last block.

    python tests/golden/make_model_rds_golden.py

Consumed by tests/test_model_rds.py (the oracle port oracle/model_rds.py on CPU; fmrx.model_chain.ModelRds on the GPU).
"""
import contextlib
import io
import os
import runpy
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
MODEL_DIR = "/root/reference/model"

from fmrx import synth  # noqa: E402

SEED, NBLK_IN = 2, 9          # the script keeps 8 blocks and never processes the last one: 7 blocks run
KEEP = ("fm_demod", "extract_rds", "pre_Pll_rds", "post_Pll", "lpf_filt_rds", "resample_rds", "rrc_rds", "symbols_I", "diff_bits", "bit_stream")
LONG_STRIDE = 4               # 15360-sample arrays are stored every 4th sample


def stub_matplotlib():
    class _Ax:
        def __getattr__(self, n):
            return lambda *a, **k: None

    plt = types.ModuleType("matplotlib.pyplot")
    plt.subplots = lambda nrows=1, **k: (_Ax(), _Ax() if nrows == 1 else tuple(_Ax() for _ in range(nrows)))
    plt.show = plt.plot = lambda *a, **k: None
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt


def run_script(raw, script="fmRDSblock.py"):
    """-> (stdout text, {block: {name: array}}, final globals)"""
    stub_matplotlib()
    if MODEL_DIR not in sys.path:
        sys.path.insert(0, MODEL_DIR)
    snaps, buf = {}, io.StringIO()

    def hooked_print(*a, **k):
        text = " ".join(str(x) for x in a)
        if text.startswith("Processing block "):
            blk = int(text.split()[-1])
            g = sys._getframe(1).f_globals
            if blk > 0:
                snaps[blk - 1] = {n: np.array(g[n], copy=True) for n in KEEP if n in g}
        print(*a, **k)

    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "model"))
        os.makedirs(os.path.join(tmp, "data"))
        raw.tofile(os.path.join(tmp, "data", "samples_rds_1029.raw"))
        os.chdir(os.path.join(tmp, "model"))
        try:
            with contextlib.redirect_stdout(buf):
                g = runpy.run_path(os.path.join(MODEL_DIR, script), init_globals={"print": hooked_print}, run_name="__main__")
        finally:
            os.chdir(cwd)
    snaps[int(g["block_count"]) - 1] = {n: np.array(g[n], copy=True) for n in KEEP if n in g}
    return buf.getvalue(), snaps, g


def main():
    raw = synth.synth_iq(NBLK_IN, 0, seed=SEED)
    text, snaps, g = run_script(raw)
    nblk = int(g["block_count"])
    d = dict(seed=SEED, nblk_in=NBLK_IN, nblk=nblk, stdout=text, final_int_offset=int(g["int_offset"]), start_pos=int(g["start_pos"]),
             printposition=int(g["printposition"]), last_position=int(g["last_position"]), prev_sync_bits=np.asarray(g["prev_sync_bits"]),
             state_Pll=np.asarray(g["state_Pll"], np.float64), state_phase=float(g["state_phase"]))
    for b in range(nblk):
        for n, v in snaps[b].items():
            if n in ("rrc_rds", "symbols_I", "diff_bits", "bit_stream", "resample_rds") or b in (0, nblk - 1):
                d[f"{n}_{b}"] = v[::LONG_STRIDE] if v.size >= 15360 else v
    np.savez_compressed(os.path.join(HERE, "model_rds.npz"), **d)
    print(text)
    print({k: np.shape(v) for k, v in d.items() if k.endswith("_0") or not k[-1].isdigit()})


if __name__ == "__main__":
    main()
