"""GPU: the ingest / egress ring (fmrx_ring_*) — a producer thread and a consumer thread around one
batch handle, bounded pinned slots, real back-pressure; results must be what the plain synchronous call gives."""
import threading
import time

import numpy as np
import pytest

import fmrx
from fmrx import synth
from oracle import Chain

pytestmark = pytest.mark.gpu


def test_ring_producer_consumer_equals_synchronous_calls():
    S, steps = 5, 7
    raw = np.stack([synth.synth_station(2 * s + 1, steps, 0) for s in range(S)])  # [S][steps * 307200]
    with fmrx.Batch(S, mode=0, profile=1, max_blocks=1) as rx:
        want = [rx.process(raw[:, k * 307200:(k + 1) * 307200]) for k in range(steps)]
    got = []
    with fmrx.Batch(S, mode=0, profile=1, max_blocks=1) as rx, fmrx.Ring(rx, n_slots=3, n_blocks=1) as ring:
        def produce():
            for k in range(steps):
                slot = ring.acquire()               # blocks while all three slots are in flight
                slot[:] = raw[:, k * 307200:(k + 1) * 307200]
                ring.commit()
            ring.close()

        t = threading.Thread(target=produce)
        t.start()
        while True:
            res = ring.next()
            if res is None:
                break
            time.sleep(0.01)                        # a slow consumer: the producer has to wait for slots
            got.append({k: v.copy() for k, v in res.items()})
            ring.release()
        t.join()
        assert ring.in_flight == 0
    assert len(got) == steps
    for k in range(steps):
        assert np.array_equal(got[k]["audio"], want[k]["audio"]), f"step {k} audio"
        assert np.array_equal(got[k]["rds_n_bits"], want[k]["rds_n_bits"]) and np.array_equal(got[k]["rds_bits"], want[k]["rds_bits"]), f"step {k} RDS bits"
    audio0 = Chain(0, 1).run(raw[0])[0]             # and the oracle, for one station end to end
    assert np.array_equal(np.concatenate([g["audio"][0].ravel() for g in got]), audio0)


def test_ring_back_pressure_timeouts_and_state_errors():
    with fmrx.Batch(2, mode=1, profile=1, max_blocks=2) as rx:
        with pytest.raises(fmrx.FmrxError):
            fmrx.Ring(rx, n_slots=1)                # a ring needs two slots
        with pytest.raises(fmrx.FmrxError):
            fmrx.Ring(rx, n_slots=3, n_blocks=3)    # more blocks per step than the handle was sized for
        with fmrx.Ring(rx, n_slots=2, n_blocks=2) as ring:
            with pytest.raises(TimeoutError):
                ring.next(timeout_ms=20)            # nothing committed yet
            with pytest.raises(fmrx.FmrxError):
                ring.commit()                       # nothing acquired
            for _ in range(2):
                ring.acquire()[:] = 128
                ring.commit()
            with pytest.raises(TimeoutError):
                ring.acquire(timeout_ms=30)         # both slots in flight and nobody consumes: back-pressure
            res = ring.next()
            assert res["audio"].shape == (2, 2, 2 * 2949) and not res["audio"].any() and "rds_bits" not in res  # silence in, mode 1: no RDS
            with pytest.raises(fmrx.FmrxError):
                ring.next()                         # the previous step was not released
            ring.release()
            ring.acquire(timeout_ms=1000)           # a slot is free again
            ring.close()                            # ... and dropped: never committed
            assert ring.next() is not None
            ring.release()
            assert ring.next() is None              # closed and drained
