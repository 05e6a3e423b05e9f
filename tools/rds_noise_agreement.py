"""RDS decisions without wide margins, at scale: S noisy synthetic stations x B blocks (white Gaussian noise ahead of the 8-bit
quantiser at the given carrier-to-noise ratio over the RF rate) through the GPU chain in each numerics setting / back end and
through the oracle on every host core; counts the decoded bits and the stations' sync-event lists that differ, and the oracle's own
bit error rate against the transmitted bits (so that the SNR means something).

    python tools/rds_noise_agreement.py [stations] [blocks] [cnr_db] > gpurun_out/<tag>_rds_noise.txt
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))

RAW = None


def oracle_station(s):
    from fmrx import synth
    from oracle import Chain

    B = RAW.shape[1] // 307200
    ch = Chain(0, 1, paths=2)
    _, _, bits, events, _ = ch.run(RAW[s])
    got = np.concatenate(bits)
    n_chips = int(np.ceil((B * 153600 - 1) / 2.4e6 * synth.CHIP_RATE)) + 2 * 4 + 2
    tx = synth.rds_bits((n_chips + 1) // 2, synth.station_params(s)["seed"])
    errs = [(int(np.count_nonzero(got[20 + max(0, -o):][:n] != tx[20 + max(0, o):][:n])), n) for o in range(-4, 5)
            for n in [min(got.size - 20 - max(0, -o), tx.size - 20 - max(0, o))]]
    e, n = min(errs)
    return s, [b.copy() for b in bits], events, ch.rds_offset, e, n


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    cnr = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
    import torch

    import fmrx
    from fmrx import synth

    global RAW
    t0 = time.time()
    raw = RAW = synth.synth_batch_torch(list(range(S)), B, 0, torch.device("cuda", 0), chunk=32, cnr_db=cnr, noise_seed=7).cpu().numpy()
    rows = (("strict, staged", fmrx.NUMERICS_STRICT, fmrx.PATH_RDS_STAGES), ("strict, symbol-rate", fmrx.NUMERICS_STRICT, 0),
            ("reference, symbol-rate (default)", fmrx.NUMERICS_REFERENCE, 0), ("fma, symbol-rate", fmrx.NUMERICS_FMA, 0))
    got = {}
    for name, numerics, extra in rows:
        with fmrx.Batch(S, mode=0, profile=1, max_blocks=B, paths=fmrx.PATH_RDS | extra, numerics=numerics) as rx:
            res = rx.process(raw)
            got[name] = (res["rds_bits"].copy(), res["rds_n_bits"].copy(), res["rds_events"].copy(), res["rds_n_events"].copy(), rx.rds_offsets().copy())
    t1 = time.time()
    with Pool(len(os.sched_getaffinity(0))) as pool:
        ref = {s: (bits, ev, off, e, n) for s, bits, ev, off, e, n in pool.imap_unordered(oracle_station, range(S), chunksize=2)}
    t2 = time.time()
    tx_e, tx_n = sum(v[3] for v in ref.values()), sum(v[4] for v in ref.values())
    print(f"CNR {cnr} dB over the 2.4 MHz RF rate, {S} stations x {B} blocks; oracle bit errors against the transmitted bits: {tx_e} / {tx_n} = {tx_e / tx_n:.3%}")
    for name, (bits, nb, ev, ne, off) in got.items():
        nbad = ntot = ev_bad = off_bad = 0
        for s in range(S):
            rb, rev, roff = ref[s][0], ref[s][1], ref[s][2]
            off_bad += int(off[s] != roff)
            for b in range(B):
                n = int(nb[s, b])
                ntot += rb[b].size
                nbad += rb[b].size if n != rb[b].size else int(np.count_nonzero(bits[s, b, :n] != rb[b]))
            g = [tuple(int(v) for v in e) for b in range(B) for e in ev[s, b, :ne[s, b]]]
            ev_bad += g != rev
        print(f"  {name:34s}: bits that differ from the oracle {nbad} / {ntot}   stations with a different sync-event list {ev_bad} / {S}   different sampling phase {off_bad}")
    print(f"  gpu + synth {t1 - t0:.1f} s, oracle {t2 - t1:.1f} s")


if __name__ == "__main__":
    main()
