"""RDS data-link / application layer (csrc/fmrx_rdsapp.cpp) against its oracle (oracle/rds_app.py) and
known-answer vectors: programmes encoded from known PI / PS / RadioText by fmrx.synth.rds_group_bits.  The layer is host
code, so everything but the last test runs without a GPU; the last one decodes the bits the GPU chain produces from a
synthetic multiplex carrying a known programme."""
import numpy as np
import pytest

import fmrx
from fmrx import synth
from oracle.rds_app import Station

PI, PS, RT = 0xC0DE, "FMRX-GPU", "Now playing: B200-native FM receive chain"


def run_both(bits, pieces=None):
    """feed the same bit stream to the library and to the oracle; returns (library groups, station dict, oracle Station)"""
    app, orc = fmrx.RdsApp(1), Station()
    bits = np.asarray(bits, np.uint8)
    got = []
    if pieces is None:
        got = app.feed_stream(bits)
    else:
        i = 0
        for n in pieces:
            got += app.feed_stream(bits[i:i + n])
            i += n
        got += app.feed_stream(bits[i:])
    ref = orc.feed(bits)
    st = app.station()
    app.close()
    return got, st, orc, ref


def same(got, ref):
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a == b, (a, b)


def check_counters(st, orc):
    assert (st["blocks_ok"], st["blocks_corrected"], st["blocks_bad"], st["sync_losses"]) == (orc.blocks_ok, orc.blocks_corrected, orc.blocks_bad, orc.sync_losses)
    assert st["ps"] == orc.ps_text and st["rt"] == orc.rt_text and st["pi"] == orc.pi and st["pty"] == orc.pty and st["tp"] == orc.tp


def test_known_programme_version_a():
    bits = synth.rds_group_bits(PI, PS, RT, pty=10, tp=1, n_bits=3 * 104 * 15)
    got, st, orc, ref = run_both(bits)
    same(got, ref)
    check_counters(st, orc)
    assert st["pi"] == PI and st["ps"] == PS and st["rt"] == RT and st["pty"] == 10 and st["tp"] == 1 and st["ps_complete"] and st["synced"]
    assert len(got) == len(bits) // 104 and all(g["corrected"] == 0 for g in got) and got[0]["bit_index"] == 0
    assert [g["type"] for g in got[:6]] == [0, 0, 0, 0, 2, 2]


def test_known_programme_version_b_uses_c_prime():
    bits = synth.rds_group_bits(0x1234, "B-GROUPS", "two chars per 2B group", version_b=True, n_bits=104 * 40)
    got, st, orc, ref = run_both(bits)
    same(got, ref)
    assert st["ps"] == "B-GROUPS" and st["rt"] == "two chars per 2B group" and all(g["version_b"] == 1 for g in got)
    assert all(g["blk"][2] == 0x1234 for g in got)  # block C' repeats the PI


def test_arbitrary_start_and_chunking_do_not_matter():
    rng = np.random.default_rng(3)
    body = synth.rds_group_bits(PI, PS, RT, n_bits=104 * 30)
    bits = np.concatenate([rng.integers(0, 2, 37, dtype=np.uint8), body[61:]])  # junk, then a stream cut in mid-block
    whole, st_w, orc, ref = run_both(bits)
    same(whole, ref)
    pieces, st_p, _, _ = run_both(bits, pieces=[1, 25, 26, 27, 80, 3, 500, 7])
    same(pieces, whole)
    assert st_w == st_p and st_w["ps"] == PS


def test_burst_errors_up_to_five_bits_are_corrected():
    bits = synth.rds_group_bits(PI, PS, RT, n_bits=104 * 40).copy()
    rng = np.random.default_rng(7)
    n_err = 0
    for blk in range(20, len(bits) // 26, 3):  # after synchronisation: every third block gets one burst
        length = int(rng.integers(1, 6))
        start = blk * 26 + int(rng.integers(0, 27 - length))
        pat = [1] + [int(v) for v in rng.integers(0, 2, max(0, length - 2))] + ([1] if length > 1 else [])
        bits[start:start + len(pat)] ^= np.array(pat, np.uint8)
        n_err += 1
    got, st, orc, ref = run_both(bits)
    same(got, ref)
    check_counters(st, orc)
    assert st["blocks_corrected"] == n_err and st["blocks_bad"] == 0 and st["ps"] == PS and st["rt"] == RT
    assert sum(g["corrected"] for g in got) == n_err


def test_noise_loses_and_regains_synchronisation():
    rng = np.random.default_rng(11)
    a = synth.rds_group_bits(PI, PS, RT, n_bits=104 * 12)
    b = synth.rds_group_bits(0x4321, "SECOND  ", "after the dropout", n_bits=104 * 20)
    bits = np.concatenate([a, rng.integers(0, 2, 26 * 40, dtype=np.uint8), b])
    got, st, orc, ref = run_both(bits)
    same(got, ref)
    check_counters(st, orc)
    assert st["sync_losses"] >= 1 and st["synced"] and st["pi"] == 0x4321 and st["rt"] == "after the dropout"


def test_random_bits_give_what_the_oracle_gives():
    bits = np.random.default_rng(5).integers(0, 2, 50000, dtype=np.uint8)
    got, st, orc, ref = run_both(bits)
    same(got, ref)
    check_counters(st, orc)


def test_batch_of_stations_and_bad_arguments():
    S, B = 5, 16
    progs = [(0x1000 + s, f"STN {s:04d}", f"station {s} radiotext") for s in range(S)]
    streams = [synth.rds_group_bits(pi, ps, rt, n_bits=B * fmrx.MAX_BITS - 11 * s) for s, (pi, ps, rt) in enumerate(progs)]
    bits = np.zeros((S, B, fmrx.MAX_BITS), np.uint8)
    n_bits = np.zeros((S, B), np.int32)
    for s, st in enumerate(streams):  # ragged: the decoder emits 75 or 76 bits per block, so blocks are not full
        i = 0
        for b in range(B):
            n = min(fmrx.MAX_BITS - (s + b) % 5, len(st) - i)
            bits[s, b, :n] = st[i:i + n]
            n_bits[s, b] = n
            i += n
    app = fmrx.RdsApp(S)
    groups = app.feed(bits, n_bits)
    for s, (pi, ps, rt) in enumerate(progs):
        orc = Station()
        ref = orc.feed(np.concatenate([bits[s, b, :n_bits[s, b]] for b in range(B)]))
        same(groups[s], ref)
        assert app.station(s)["pi"] == pi and app.station(s)["ps"] == ps
    with pytest.raises(fmrx.FmrxError):
        app.station(S)
    n_bits[0, 0] = fmrx.MAX_BITS + 1
    with pytest.raises(fmrx.FmrxError):
        app.feed(bits, n_bits)
    app.reset()
    assert app.station(0)["pi"] == -1 and app.station(0)["ps"] == "________"


@pytest.mark.gpu
def test_programme_decoded_from_the_gpu_chain():
    """a synthetic multiplex carrying a known programme -> the CUDA chain -> bits -> PI / PS / RadioText"""
    nblk = 40  # 2.56 s: 3040 bits, 29 groups
    raw = synth.synth_iq(nblk, 0, seed=1, rds_payload=lambda n: synth.rds_group_bits(PI, PS, RT, pty=5, n_bits=n))
    with fmrx.Batch(1, mode=0, profile=1, max_blocks=nblk) as rx:
        res = rx.process(raw)
    app = fmrx.RdsApp(1)
    groups = app.feed(res["rds_bits"], res["rds_n_bits"])[0]
    st = app.station()
    print(st)
    orc = Station()
    same(groups, orc.feed(np.concatenate([res["rds_bits"][0, b, :res["rds_n_bits"][0, b]] for b in range(nblk)])))
    assert st["pi"] == PI and st["ps"] == PS and st["pty"] == 5 and st["rt"] == RT and st["blocks_bad"] == 0 and len(groups) >= 25
