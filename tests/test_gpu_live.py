"""GPU: the live path at wall-clock rate.  The reference is used as `rtl_sdr | fm_radio | aplay`
(/root/reference/src/fm_radio.cpp:86-138: blocking reads pace the input, blocking writes the output).  Two paced runs:

  * the `fm_radio` executable fed 307200 bytes every 64 ms (2.4 Msps) on stdin: every block must come out on stdout, in
    order, bit-identical to the oracle, with a per-block latency (last input byte written -> block readable on stdout) far
    inside the 64 ms budget;
  * the library's ingest ring in front of a 4096-station batch, one step committed every 64 ms (4096 x 2.4 Msps = 9.8 Gsps
    offered, 1.26 GB per step over PCIe): no step may find the ring full (that would be a drop on a live receiver) and the
    commit -> results-on-the-host latency must stay inside the block period.
"""
import os
import subprocess
import threading
import time

import numpy as np
import pytest

import fmrx
from fmrx import synth
from oracle import Chain

pytestmark = pytest.mark.gpu
BLOCK, PERIOD = 307200, 0.064


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(round(q * (len(v) - 1))))]


def test_cli_paced_stdin_is_real_time():
    warm, paced = 4, 40
    raw = synth.synth_iq(warm + paced, 0, seed=12)
    p = subprocess.Popen([fmrx.CLI_PATH, "--quiet"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, bufsize=0)
    out_bytes = 2 * 3072 * 2
    got, t_out = [], []

    def reader():
        while True:
            buf = b""
            while len(buf) < out_bytes:
                chunk = p.stdout.read(out_bytes - len(buf))
                if not chunk:
                    return
                buf += chunk
            t_out.append(time.perf_counter())
            got.append(buf)

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    # start-up (CUDA context, handle creation) is not part of the steady state: push a few blocks through first
    for k in range(warm):
        p.stdin.write(raw[k * BLOCK:(k + 1) * BLOCK].tobytes())
    deadline = time.perf_counter() + 120
    while len(got) < warm and time.perf_counter() < deadline:
        time.sleep(0.01)
    assert len(got) == warm, "the executable did not produce the warm-up blocks"
    t_in, t0 = [], time.perf_counter()
    for k in range(paced):
        target = t0 + k * PERIOD
        while time.perf_counter() < target:
            time.sleep(0.0005)
        p.stdin.write(raw[(warm + k) * BLOCK:(warm + k + 1) * BLOCK].tobytes())
        t_in.append(time.perf_counter())
    p.stdin.close()
    th.join(timeout=60)
    assert p.wait(timeout=60) == 0
    assert len(got) == warm + paced, f"{warm + paced - len(got)} blocks never came out"
    lat = [(t_out[warm + k] - t_in[k]) * 1e3 for k in range(paced)]
    span = t_in[-1] - t_in[0]
    print(f"fm_radio paced at {PERIOD * 1e3:.0f} ms per block over {span:.2f} s: latency ms p50 {pct(lat, 0.5):.2f}  p90 {pct(lat, 0.9):.2f}  p99 {pct(lat, 0.99):.2f}  max {max(lat):.2f}")
    assert abs(span - (paced - 1) * PERIOD) < 0.5, "the feeder itself did not keep the cadence"
    # measured: 4 ms.  The bars leave room for a noisy host (the feeder and the reader are Python threads): median inside half a block
    # period, nothing later than two periods
    assert pct(lat, 0.5) < 0.5 * PERIOD * 1e3 and max(lat) < 2 * PERIOD * 1e3, "a block took longer than its own duration"
    audio = np.frombuffer(b"".join(got), dtype=np.int16)
    ref = Chain(0, 0).run(raw)[0]
    assert np.array_equal(audio, ref), "paced output differs from the oracle (binary profile)"


@pytest.mark.parametrize("stations", [int(os.environ.get("FMRX_LIVE_STATIONS", "4096"))])
def test_ring_paced_full_batch_is_real_time(stations):
    import torch

    S, slots, steps = stations, 4, 40
    dev = torch.device("cuda", 0)
    d_iq = synth.synth_batch_torch(range(min(S, 256)), 1, 0, dev, chunk=64)
    block = d_iq.cpu().numpy()
    del d_iq
    torch.cuda.empty_cache()
    with fmrx.Batch(S, mode=0, profile=fmrx.PROFILE_INTENT, max_blocks=1) as rx, fmrx.Ring(rx, n_slots=slots, n_blocks=1) as ring:
        # what the SDR's DMA would have written: every slot holds a block for every station before the clock starts
        filled = []
        for _ in range(slots):
            buf = ring.acquire()
            for lo in range(0, S, block.shape[0]):
                n = min(block.shape[0], S - lo)
                buf[lo:lo + n] = block[:n]
            filled.append(buf)
            ring.commit()
        for _ in range(slots):
            assert ring.next(timeout_ms=60000) is not None
            ring.release()
        t_commit, t_done, drops, depth = [], [], [0], []

        def consumer():
            for _ in range(steps):
                res = ring.next(timeout_ms=60000)
                t_done.append(time.perf_counter())
                assert res is not None and res["rds_n_bits"].shape == (S, 1)
                ring.release()

        th = threading.Thread(target=consumer)
        th.start()
        t0 = time.perf_counter()
        for k in range(steps):
            target = t0 + k * PERIOD
            while time.perf_counter() < target:
                time.sleep(0.0005)
            try:
                ring.acquire(timeout_ms=0)      # a live producer cannot wait: a full ring is a dropped step
            except TimeoutError:
                drops[0] += 1
                ring.acquire(timeout_ms=-1)
            depth.append(ring.in_flight)
            ring.commit()
            t_commit.append(time.perf_counter())
        th.join(timeout=120)
        ring.close()
        assert ring.next(timeout_ms=1000) is None
    lat = [(b - a) * 1e3 for a, b in zip(t_commit, t_done)]
    offered = S * 153600 / PERIOD / 1e9
    print(f"ring, {S} stations, one step per {PERIOD * 1e3:.0f} ms ({offered:.2f} Gsps offered): commit -> results latency ms p50 {pct(lat, 0.5):.1f}  p90 {pct(lat, 0.9):.1f}  "
          f"p99 {pct(lat, 0.99):.1f}  max {max(lat):.1f}; steps in flight at commit: max {max(depth)}; drops {drops[0]}")
    assert drops[0] == 0 and len(t_done) == steps
    assert max(depth) <= 2, "steps piled up in the ring: the pipeline does not keep up with the offered rate"
    # measured: 33.6 ms median, 37.1 ms maximum (23 ms of it is the 1.26 GB host-to-device copy); median inside the block period,
    # no step later than two periods
    assert pct(lat, 0.5) < PERIOD * 1e3 and max(lat) < 2 * PERIOD * 1e3
