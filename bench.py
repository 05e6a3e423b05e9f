#!/usr/bin/env python
"""bench.py — throughput of the batched FM receive chain (BASELINE.json config 5) on N B200s of one node.

    python bench.py [--gpus N --steps K --warmup W]                      (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference [--gpus N --steps K --warmup W]     (the reference's own CPU code, rank 0 only)

Workload (config.workload = "batch4096_mode0_stereo_rds"): `--stations` (4096) independent synthetic FM stations PER
GPU (weak scaling; stations are independent, so ranks share nothing and there is no collective on the data path),
mode 0 = 2.4 Msps 8-bit IQ -> mono + stereo + RDS, `intent` profile (stereo computed in every block), one 307200-byte
block (64 ms of signal) per station per step, filter / PLL / decoder state carried from step to step.  The same
synthesised block is replayed every step (the arithmetic is data-independent); the input of one step is 1.26 GB, ten
times the L2, so no L2 flush is needed.

metric  : complex IQ samples consumed per second, whole job, in Msps;  real-time streams = value / 2.4.
value   : inputs already resident in HBM, timed with CUDA events on the library's compute stream, max over ranks.
e2e     : the same K steps through fmrx_batch_process with pinned HOST buffers (H2D of the IQ bytes and D2H of audio +
          RDS results inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))

BLOCK_BYTES = 307200
BLOCK_IQ = 153600
NIF, NAUD, NRDS, NT = 15360, 3072, 3648, 151
# algorithmic MACs per station-block and whether the stage keeps the reference's two roundings per tap
STAGE_MACS = {
    "frontend": (NIF * NT * 2, True), "mono": (NAUD * NT, True), "pilot_bpf": (NIF * NT, True), "stereo_bpf": (NIF * NT, True),
    "rds_bpf": (NIF * NT, False), "rds_sq_bpf": (NIF * NT, False), "stereo_lpf": (NAUD * NT, True), "rds_mix_lpf": (NIF * NT, False),
    "rds_resample": (NRDS * NT, False), "rds_rrc": (NRDS * NT, False),
    # the symbol-rate RDS back end (csrc/fmrx_rdsfast.cu) does the WORK it does, not the reference's 3.4 M MACs: head 1008 + 240 + 34
    # outputs x 151, 142 interior symbols x 933, state 150 x 301 + 8 x 151
    "rds_symbols": ((1008 + 240 + 34) * NT + 142 * 933 + 150 * 301 + 8 * NT, False),
}
# algorithmic HBM bytes per station-block for the kernel split actually used (u8 in, fp32 intermediates, int16 out)
STAGE_BYTES = {
    "frontend": BLOCK_BYTES + 4 * NIF, "mono": 4 * NIF + 4 * NAUD, "pilot_bpf": 8 * NIF, "stereo_bpf": 8 * NIF, "rds_bpf": 8 * NIF,
    # pll: two loops, each reads its input and the signal it is mixed with and writes the NCO and the product: 4 x 4 B per loop-sample
    "rds_sq_bpf": 8 * NIF, "pll": 32 * NIF, "stereo_lpf": 4 * NIF + 4 * NAUD, "combine": 8 * NAUD + 4 * NAUD, "rds_mix_lpf": 8 * NIF,
    "rds_resample": 4 * NIF + 4 * NRDS, "rds_rrc": 8 * NRDS, "rds_decode": 4 * 152 + 80 + 8 + 2 * 4 * 160, "rds_symbols": 4 * NIF + 4 * 152,
}


PROFILE_TAG = "r5"  # profiles/<tag>_ncu_kernels.csv (tools/ncu_trim.py) and profiles/<tag>_sass_mix.json (tools/sass_mix.py)
# which kernel launches of one chain step make up a bench stage (names as tools/ncu_trim.py shortens them; in launch order)
STAGE_KERNELS = {
    "frontend": ("frontend_stream4_kernel<1>", "frontend_edge_kernel<1>", "iq_state_kernel<1>"),
    "pll": ("pll_kernel",), "combine": ("combine4_kernel", "combine_kernel"), "rds_decode": ("rds_decode_kernel",),
    "rds_symbols": ("rds_symbol_kernel",), "bpf_fused": ("fir151_multi_kernel",),
}


def ncu_profile():
    """{kernel: [rows]} from the tracked ncu summary, or {} when it is missing.  One capture = one chain step of 4096 stations."""
    import csv

    path = os.path.join(ROOT, "profiles", f"{PROFILE_TAG}_ncu_kernels.csv")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r4_ncu_kernels.csv")  # the previous capture (same kernels except the PLL step variant)
    out = {}
    if os.path.exists(path):
        for r in csv.DictReader(open(path)):
            out.setdefault(r["kernel"], []).append(r)
    return out, os.path.relpath(path, ROOT)


def stage_rows(prof, stage, fir_order):
    """the launches that make up `stage` in the committed capture (4096 stations x 1 block)"""
    if stage in STAGE_KERNELS:
        return [r for k, v in prof.items() if any(k.startswith(p) for p in STAGE_KERNELS[stage]) for r in v]
    if stage in fir_order:  # the FIR stages share one kernel template: identify the launch by its position among the fir151_kernel launches
        firs = sorted((r for k, v in prof.items() if k.startswith("fir151_kernel") for r in v), key=lambda r: int(r["launch"]))
        i = fir_order.index(stage)
        return [firs[i]] if i < len(firs) else []
    return []


def stage_traffic(prof, stage, fir_order):
    """dram read + write bytes of those launches, or None"""
    rows = stage_rows(prof, stage, fir_order)
    return sum(float(r["dram_read_bytes"]) + float(r["dram_write_bytes"]) for r in rows) if rows else None


def stage_pipes(prof, stage, fir_order):
    """pipe utilisation of the stage's longest launch in the committed capture (ncu, % of peak while the SM is active): what the kernel is bound by"""
    rows = stage_rows(prof, stage, fir_order)
    if not rows:
        return None
    r = max(rows, key=lambda x: float(x["duration_ms"]))
    keys = (("issue_active_pct", "issue"), ("pipe_fma_pct", "fma"), ("pipe_alu_pct", "alu"), ("pipe_fp64_pct", "fp64"), ("pipe_xu_pct", "xu"), ("pipe_lsu_pct", "lsu"))
    out = {lab: round(float(r[k]), 1) for k, lab in keys if r.get(k) not in (None, "")}
    out["busiest"] = max((k for k in out if k != "issue"), key=lambda k: out[k]) if len(out) > 1 else None
    return out


def sass_mix():
    path = os.path.join(ROOT, "profiles", f"{PROFILE_TAG}_sass_mix.json")
    return (json.load(open(path)) if os.path.exists(path) else {}), os.path.relpath(path, ROOT)


def python_models_baseline(blocks=2):
    """the reference's Python block models on one host core, seconds per 307200-byte block of signal.  Where the
    reference's model directory is mounted (the authoring container) the UNMODIFIED scripts are timed through runpy (kind
    "reference"); on the GPU box, where it does not exist, the oracle's statement-by-statement port of the same loops is timed
    (kind "port": oracle/model_rds.py, pinned bit for bit to the scripts by tests/test_model_rds.py)."""
    from fmrx import synth

    out = {"cores": 1}
    model_dir = "/root/reference/model"
    raw = synth.synth_iq(blocks + 1, 0, seed=2)
    if os.path.isdir(model_dir):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from make_model_rds_golden import run_script

        t0 = time.perf_counter()
        _, _, g = run_script(raw[:(blocks + 1) * BLOCK_BYTES])
        dt = time.perf_counter() - t0
        out["fmRDSblock"] = {"s_per_block": round(dt / max(1, int(g["block_count"])), 3), "blocks": int(g["block_count"]), "kind": "reference"}
    from oracle.model_rds import ModelMonoPort, ModelRdsPort

    port = ModelRdsPort()
    t0 = time.perf_counter()
    for b in range(blocks):
        port.block(raw[b * BLOCK_BYTES:(b + 1) * BLOCK_BYTES])
    out["fmRDSblock_port"] = {"s_per_block": round((time.perf_counter() - t0) / blocks, 3), "blocks": blocks, "kind": "port"}
    iq = ((raw[:blocks * BLOCK_BYTES].astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    mono = ModelMonoPort()
    t0 = time.perf_counter()
    for b in range(blocks * 3):  # the mono script's block is 102400 values: three per 307200
        mono.block(iq[b * 102400:(b + 1) * 102400])
    out["fmMonoBlock_port"] = {"s_per_block": round((time.perf_counter() - t0) / blocks, 3), "blocks": blocks, "kind": "port",
                               "note": "mono + stereo loop of fmMonoBlock.py, three of its 102400-value blocks per 307200-byte block"}
    for k in ("fmRDSblock", "fmRDSblock_port", "fmMonoBlock_port"):
        if k in out:
            out[k]["msps"] = round(BLOCK_IQ / out[k]["s_per_block"] / 1e6, 3)
    return out


def env_int(name, default):
    return int(os.environ.get(name, default))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation, timed on the host cores (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------------------
_CPU_SAMPLES = {}


def _cpu_sample(n_blocks, mode):
    if (n_blocks, mode) not in _CPU_SAMPLES:
        from fmrx import synth

        _CPU_SAMPLES[(n_blocks, mode)] = synth.synth_iq(n_blocks, mode, seed=1)
    return _CPU_SAMPLES[(n_blocks, mode)]


def cpu_reference_run(n_blocks, procs, mode=0):
    """P independent processes of the reference path over `n_blocks` blocks each.  Uses the unmodified reference
    executable (oracle/_ref/fm_radio, kind "reference") when it was built, else the oracle port's chain (kind "port").
    Returns (Msps aggregate, kind, seconds)."""
    raw = _cpu_sample(n_blocks, mode)  # synthesised once per (length, mode): the reference arm calls this once per step
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "fm_radio")
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(tmpdir, f"fmrx_bench_{os.getpid()}.raw")
    raw.tofile(path)
    try:
        if os.path.exists(ref_bin):
            kind = "reference"
            cmd = [ref_bin] + (["1"] if mode == 1 else [])
            t0 = time.perf_counter()
            ps = [subprocess.Popen(cmd, stdin=open(path, "rb"), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(procs)]
            for p in ps:
                p.wait()
            dt = time.perf_counter() - t0
        else:
            kind = "port"
            code = ("import sys,numpy as np;sys.path.insert(0,%r);from oracle import Chain;c=Chain(%d,1);"
                    "raw=np.fromfile(%r,np.uint8);[c.block(raw[b*307200:(b+1)*307200]) for b in range(%d)]" % (ROOT, mode, path, n_blocks))
            subprocess.run([sys.executable, "-c", "import sys;sys.path.insert(0,%r);from oracle import load_port;load_port()" % ROOT], check=True)
            t0 = time.perf_counter()
            ps = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for _ in range(procs)]
            for p in ps:
                p.wait()
            dt = time.perf_counter() - t0
    finally:
        os.unlink(path)
    return procs * n_blocks * BLOCK_IQ / dt / 1e6, kind, dt


def cpu_stage_times(reps=3):
    """the reference's own functions (oracle/_ref/libfmref.so: the unmodified reference objects behind a
    C shim; the oracle port where that was not built) on ONE host core, one mode-0 block, per stage of the chain --
    milliseconds per block, the figure to hold against the GPU's stage time / stations.  Returns (dict, kind)."""
    from fmrx import synth
    from oracle.ref import ref_available

    if ref_available():
        from oracle.ref import Ref
        r, kind = Ref(), "reference"
    else:
        from oracle import Port
        r, kind = Port(), "port"
    F = np.float32
    raw = synth.synth_iq(1, 0, seed=1)
    z = lambda n=150: np.zeros(n, F)  # noqa: E731
    h_rf, h_a = r.lpf(2.4e6, 1e5, 151), r.lpf(240e3, 16e3, 151)
    h_p, h_s = r.bpf(18.5e3, 19.5e3, 240e3, 151), r.bpf(22e3, 54e3, 240e3, 151)
    h_r, h_q, h_l = r.bpf(54e3, 60e3, 240e3, 151), r.bpf(113.5e3, 114.5e3, 240e3, 151), r.lpf(240e3, 3e3, 151)
    h_anti, h_rrc = r.lpf(240e3 * 19, 57000 // 2, 151 * 19), r.rrc(57e3, 151)
    pll0 = lambda: np.array([0, 0, 1, 0, 0, 1], F)  # noqa: E731
    phase = F(float(F(np.pi / 3.3 - np.pi / 1.5)) - np.pi / 1.4)
    out = {}

    def timed(name, fn):
        best, res = 1e9, None
        for _ in range(reps):
            t0 = time.perf_counter()
            res = fn()
            best = min(best, time.perf_counter() - t0)
        out[name] = round(best * 1e3, 3)
        return res

    iq = timed("unpack", lambda: r.unpack(raw))
    i_, q_ = iq[0::2].copy(), iq[1::2].copy()
    yi, yq = timed("frontend_fir", lambda: r.fir_decim_iq(i_, q_, h_rf, z(), z(), 10))
    demod = timed("discriminator", lambda: r.demod(yi, yq))
    out["frontend"] = round(out["unpack"] + out["frontend_fir"] + out["discriminator"], 3)
    timed("mono", lambda: r.fir_decim(demod, h_a, z(), 5))
    pilot = timed("pilot_bpf", lambda: r.fir_decim(demod, h_p, z(), 1))
    sb = timed("stereo_bpf", lambda: r.fir_decim(demod, h_s, z(), 1))
    nco = timed("pll_pilot", lambda: r.pll(pilot, 19e3, 240e3, 2.0, 0.0, 0.01, pll0()))
    timed("stereo_lpf", lambda: r.fir_decim((sb * nco[:sb.size]).astype(F), h_a, z(), 5))
    rb = timed("rds_bpf", lambda: r.fir_decim(demod, h_r, z(), 1))
    _, rnco = timed("rds_sq_bpf_and_pll", lambda: r.pll_combine(rb, h_q, z(), 114e3, 240e3, 0.5, phase, 0.001, pll0()))
    lp = timed("rds_mix_lpf", lambda: r.fir_mixer(rnco, rb, h_l, z()))
    res_fn = getattr(r, "resample_rds", None)
    rr = timed("rds_resample", (lambda: res_fn(lp, h_anti, z(151 * 19 - 1), 80, 19)) if res_fn else (lambda: r.resample(lp, h_anti, z(151 * 19 - 1), 80, 19, True)))
    timed("rds_rrc", lambda: r.fir_decim(rr, h_rrc, z(), 1))
    out["sum"] = round(sum(v for k, v in out.items() if k not in ("unpack", "frontend_fir", "discriminator")), 3)
    return out, kind


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(index):
    """Pins this rank's threads to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it
    allocates (first touch) and the H2D copies out of them stay on that node's memory controller.  Returns the node or None."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
        dev_id = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def pick_gpu(rank, world, local_rank):
    """Which visible GPU this rank drives.  With as many ranks as GPUs (or one rank) it is LOCAL_RANK.  With FEWER ranks than visible
    GPUs the choice matters for the end-to-end figure: on the 8-GPU boxes of this pool four of the GPUs share one host-side limit
    (GPUs 0-3 together ingest 115 GB/s from page-locked memory, GPUs 4-7 together 214 GB/s, any pair 111 GB/s: profiles/r4_h2d_probe8.json;
    the virtual PCI tree is flat and reports no NUMA node, so sysfs does not say which).  Rank 0 therefore MEASURES it before anything
    else runs -- concurrent host-to-device copies only, no kernels -- and grows the set greedily: starting from GPU 0, add the GPU that
    gives the largest aggregate rate together with those already chosen (ties to the lowest index).  The map is handed to the other
    ranks through a file in /dev/shm keyed by the launcher's pid (the ranks have no process group yet).  FMRX_GPU_MAP=identity disables it."""
    import torch

    n_vis = torch.cuda.device_count()
    info = {"visible_gpus": n_vis, "policy": "identity", "chosen": None}
    if world == 1 or n_vis <= world or os.environ.get("FMRX_GPU_MAP", "") == "identity":
        return local_rank, info
    path = f"/dev/shm/fmrx_gpumap_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}.json"
    if rank == 0:
        nbytes, reps = 256 << 20, 3
        dev, host, streams = [], [], []
        for i in range(n_vis):
            torch.cuda.set_device(i)
            dev.append(torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}"))
            host.append(torch.empty(nbytes, dtype=torch.uint8).pin_memory())
            streams.append(torch.cuda.Stream(device=i))

        def rate(group):
            for rep in range(reps + 1):
                if rep == 1:
                    for i in group:
                        streams[i].synchronize()
                    t0 = time.perf_counter()
                for i in group:
                    with torch.cuda.stream(streams[i]):
                        dev[i].copy_(host[i], non_blocking=True)
            for i in group:
                streams[i].synchronize()
            return len(group) * reps * nbytes / (time.perf_counter() - t0) / 1e9

        chosen, trace = [0], []
        while len(chosen) < world:
            cand = {c: rate(chosen + [c]) for c in range(n_vis) if c not in chosen}
            best = max(cand.values())
            pick = min(c for c, v in cand.items() if v >= 0.97 * best)
            trace.append({"with": pick, "gbs": round(cand[pick], 1), "worst_alternative_gbs": round(min(cand.values()), 1)})
            chosen.append(pick)
        info.update({"policy": "greedy on measured concurrent H2D rate (rank 0, copies only)", "chosen": chosen, "trace": trace,
                     "identity_gbs": round(rate(list(range(world))), 1), "chosen_gbs": round(rate(chosen), 1)})
        del dev, host, streams
        torch.cuda.empty_cache()
        tmp = path + ".tmp"
        json.dump(info, open(tmp, "w"))
        os.replace(tmp, path)
    else:
        t0 = time.time()
        while not os.path.exists(path) and time.time() - t0 < 300:
            time.sleep(0.05)
        if os.path.exists(path):
            info = json.load(open(path))
        else:
            info["policy"] = "identity (no map from rank 0 within 300 s)"
    chosen = info.get("chosen") or list(range(world))
    return chosen[local_rank % len(chosen)], info


class ShmBarrier:
    """Host barrier between the ranks of one node through a few words of POSIX shared memory: every rank owns one slot and writes its
    epoch there, a barrier is over when every slot has reached the epoch.  One writer per slot, so no atomics; a spin of a few
    microseconds -- the ingest arbitration crosses two of these per step and a gloo barrier (a TCP ring) cost it milliseconds."""

    def __init__(self, rank, world):
        from multiprocessing import shared_memory

        self.rank, self.world, self.epoch = rank, world, 0
        name = f"fmrx_bar_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=8 * world)
            self.shm.buf[:8 * world] = bytes(8 * world)
        else:
            t0 = time.time()
            while True:
                try:
                    self.shm = shared_memory.SharedMemory(name=name)
                    if self.shm.size >= 8 * world:
                        break
                except FileNotFoundError:
                    pass
                if time.time() - t0 > 120:
                    raise RuntimeError("no shared-memory barrier from rank 0")
                time.sleep(0.01)
            try:  # the creator unlinks it; keep this process's resource tracker from doing (and announcing) the same at exit
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.slots = np.ndarray((world,), dtype=np.int64, buffer=self.shm.buf)

    def wait(self):
        self.epoch += 1
        self.slots[self.rank] = self.epoch
        while int(self.slots.min()) < self.epoch:
            pass

    def close(self):
        self.slots = None
        self.shm.close()
        if self.rank == 0:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass


def run_reference_arm(args, rank):
    if rank != 0:
        return
    procs, nblk = host_cores(), args.cpu_blocks
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_run(max(2, nblk // 4), procs)
    t_total, units, kind = 0.0, 0, "reference"
    for _ in range(args.steps):
        msps, kind, dt = cpu_reference_run(nblk, procs)
        t_total += dt
        units += procs * nblk * BLOCK_IQ
    value = units / t_total / 1e6
    sample = f"{procs} concurrent processes x {nblk} blocks (mode 0: mono+stereo+RDS) per step, input from tmpfs, stdout to /dev/null"
    line = {
        "impl": "reference", "metric": "IQ Msps (complex samples/s, whole job)", "value": round(value, 3), "unit": "Msps", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_total / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch4096_mode0_stereo_rds", "stations_per_gpu": args.stations, "blocks_per_step": args.blocks, "mode": 0, "paths": "mono+stereo+rds",
                   "profile": "the shipped executable (= `binary` profile: its stereo branch is dead after block 0, SURVEY Q7)",
                   "reference_sample": sample, "realtime_streams": round(value / 2.4, 2)},
        "cpu_baseline": {"value": round(value, 3), "unit": "Msps", "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Msps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_fmrx_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import fmrx
    from fmrx import shard, synth

    gpu_index, gpu_map = pick_gpu(rank, world, local_rank)
    local_rank = gpu_index  # from here on: the CUDA device this rank drives
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # pinned buffers are first-touched on the GPU's own node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()  # every rank has read the GPU map by now
        if rank == 0:
            try:
                os.unlink(f"/dev/shm/fmrx_gpumap_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}.json")
            except OSError:
                pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    S, B = args.stations, args.blocks
    stations = shard.weak_range(S, rank)  # weak scaling: every rank owns S stations, numbered globally; no collective on the data path
    t0 = time.perf_counter()
    d_iq = synth.synth_batch_torch(stations, B, 0, dev, chunk=64)
    torch.cuda.synchronize()
    t_synth = time.perf_counter() - t0

    numerics = {"reference": fmrx.NUMERICS_REFERENCE, "strict": fmrx.NUMERICS_STRICT, "fma": fmrx.NUMERICS_FMA}[args.numerics]
    rx = fmrx.Batch(S, mode=0, profile=fmrx.PROFILE_INTENT, max_blocks=B, device=local_rank, numerics=numerics)
    na = rx.n_audio
    d_audio = torch.empty((S, B, 2 * na), dtype=torch.int16, device=dev)
    d_bits = torch.zeros((S, B, fmrx.MAX_BITS), dtype=torch.uint8, device=dev)
    d_nbits = torch.zeros((S, B), dtype=torch.int32, device=dev)
    d_ev = torch.zeros((S, B, fmrx.MAX_EVENTS, 4), dtype=torch.int32, device=dev)
    d_nev = torch.zeros((S, B), dtype=torch.int32, device=dev)

    def ptr(t, typ):
        return C.cast(C.c_void_p(t.data_ptr()), typ)

    dout = fmrx.Outputs(ptr(d_audio, fmrx.i16p), None, ptr(d_bits, fmrx.u8p), ptr(d_nbits, fmrx.i32p), ptr(d_ev, fmrx.evp), ptr(d_nev, fmrx.i32p))

    # ---- parity spot check on the first block of a few stations, against the oracle on the very same bytes
    parity = "skipped"
    if rank == 0 and not args.no_check:
        from oracle import Chain

        rx.process_device(d_iq.data_ptr(), B, dout)
        rx.sync()
        for s in sorted({0, 1, min(63, S - 1), S - 1}):
            raw = d_iq[s].cpu().numpy()
            audio, _, bits, events, _ = Chain(0, 1).run(raw)
            assert np.array_equal(d_audio[s].cpu().numpy().ravel(), audio), f"station {s}: audio differs from the oracle"
            got = np.concatenate([d_bits[s, b, :int(d_nbits[s, b])].cpu().numpy() for b in range(B)])
            assert np.array_equal(got, np.concatenate(bits)), f"station {s}: RDS bits differ from the oracle"
        parity = "audio bit-exact, RDS bits equal vs oracle on stations {0,1,63,S-1}"
    rx.reset()

    # the device path is a 3-stream pipeline: time from the first stream's start to the last stream's end
    stream_first = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream_phase(rx.h, 0), device=dev)
    stream_last = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream(rx.h), device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- device-resident throughput
    for _ in range(args.warmup):
        rx.process_device(d_iq.data_ptr(), B, dout)
    rx.sync()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = rx.launches
    ev0.record(stream_first)
    for _ in range(args.steps):
        rx.process_device(d_iq.data_ptr(), B, dout)
    ev1.record(stream_last)
    rx.sync()
    barrier()
    ms_dev = max_over_ranks(ev0.elapsed_time(ev1))
    launches = rx.launches - l0
    clocks = sampler.stop() if sampler else None
    # ---- per-stage times: a separate pass of the same K steps with the three phases serialised on one stream, every
    # stage bracketed by CUDA events on that stream (inside the pipelined run above, stages of different steps overlap
    # and an event pair would time the contention, not the kernel)
    rx.profile(True)
    for _ in range(args.steps):
        rx.process_device(d_iq.data_ptr(), B, dout)
    stage = rx.stage_times()
    rx.profile(False)
    units = world * S * B * BLOCK_IQ * args.steps
    value = units / (ms_dev * 1e-3) / 1e6
    # ---- the PLL kernel as it runs in the timed region: pipelined, on its own SM partition, beside the filters of the neighbouring
    # steps.  Timeline mode keeps the three-stream pipeline and brackets every stage with events on the stream it runs on.
    n_tl = min(args.steps, 30)
    rx.profile(2)
    for _ in range(n_tl):
        rx.process_device(d_iq.data_ptr(), B, dout)
    tl = rx.timeline()
    rx.profile(False)
    pll_tl = [t1 - t0 for name, t0, t1 in tl if name == "pll"][3:]  # the first steps fill the pipeline
    pll_pipelined_ms = sum(pll_tl) / len(pll_tl) if pll_tl else None

    pll_sms, filter_sms = rx.partition()
    if args.device_only or args.skip_e2e:
        if rank == 0:
            print(json.dumps({"device_only": True, "value": round(value, 1), "unit": "Msps", "ms_per_step": round(ms_dev / args.steps, 4), "pll_sms": pll_sms,
                              "filter_sms": filter_sms, "pll_pipelined_ms": None if pll_pipelined_ms is None else round(pll_pipelined_ms, 4),
                              "stages_ms": {k: round(v[0] / args.steps, 4) for k, v in stage.items() if v[1]}}), flush=True)
        return

    # ---- end to end: pinned host buffers in and out, every step's copies inside the timed region.  The caller's loop is
    # the streaming one a receiver runs: submit step k (H2D of its 1.26 GB, the kernels, D2H of its results), then wait
    # for step k-1 and consume it -- at most two steps in flight, two pinned input buffers and two result sets.
    h_iq = [torch.empty((S, B * BLOCK_BYTES), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    for t in h_iq:
        t.copy_(d_iq)
    hsets = []
    for _ in range(2):
        hs = {"audio": torch.empty((S, B, 2 * na), dtype=torch.int16, pin_memory=True), "bits": torch.zeros((S, B, fmrx.MAX_BITS), dtype=torch.uint8, pin_memory=True),
              "nbits": torch.zeros((S, B), dtype=torch.int32, pin_memory=True), "ev": torch.zeros((S, B, fmrx.MAX_EVENTS, 4), dtype=torch.int32, pin_memory=True),
              "nev": torch.zeros((S, B), dtype=torch.int32, pin_memory=True)}
        hs["out"] = fmrx.Outputs(ptr(hs["audio"], fmrx.i16p), None, ptr(hs["bits"], fmrx.u8p), ptr(hs["nbits"], fmrx.i32p), ptr(hs["ev"], fmrx.evp), ptr(hs["nev"], fmrx.i32p))
        hsets.append(hs)
    h2d = h_iq[0].numel()
    d2h = sum(hsets[0][k].numel() * hsets[0][k].element_size() for k in ("audio", "bits", "nbits", "ev", "nev"))

    def e2e_steps(n):
        tickets = []
        for k in range(n):
            tickets.append(rx.submit(h_iq[k % 2].data_ptr(), B, hsets[k % 2]["out"]))
            if k >= 1:
                rx.wait(tickets[k - 1])  # step k-1's results are on the host now
        rx.wait(tickets[-1])

    rx.reset()
    e2e_steps(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize()
    s_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = units / s_e2e / 1e6
    # the same through the synchronous single call (returns when its own results have landed: nothing overlaps across calls)
    rx.reset()
    rx.process_into(h_iq[0].data_ptr(), B, hsets[0]["out"])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rx.process_into(h_iq[0].data_ptr(), B, hsets[0]["out"])
    torch.cuda.synchronize()
    s_sync = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_sync_value = units / s_sync / 1e6
    if rank == 0 and not args.no_check:
        assert (hsets[0]["audio"].numpy() != 0).any(), "end-to-end path produced no audio"
    # what bounds the end-to-end figure: the host->device link.  The same pinned buffer copied alone, timed on the device.
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_iq.copy_(h_iq[0], non_blocking=True)
    torch.cuda.synchronize()
    c0.record()
    for _ in range(5):
        d_iq.copy_(h_iq[0], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    link_gbs = 5 * h2d / (c0.elapsed_time(c1) * 1e-3) / 1e9

    # the same with every rank copying at once, between barriers: the box's ceiling for this many GPUs ingesting together
    def group_copy_seconds(active):
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(5):
                d_iq.copy_(h_iq[0], non_blocking=True)
            torch.cuda.synchronize()
        t = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return t

    all_gbs = world * 5 * h2d / group_copy_seconds(True) / 1e9
    # Ingest arbitration.  On this pool's 8-GPU boxes the host feeds four GPUs at once faster than eight (214 GB/s
    # against 186), so when that is what the copies alone show, the two halves of the ranks (even / odd) take turns on the link:
    # a host-side barrier, one half submits its step and waits for its host-to-device copy, barrier, the other half.  Kernels and
    # result copies of either half run whenever they are ready; only the big input copies are serialised between the halves.
    ingest = {"mode": "free-running", "free_running_msps": round(e2e_value, 1), "all_ranks_copying_gbs": round(all_gbs, 1)}
    if world >= 4:
        my_half = rank % 2
        t_turns = group_copy_seconds(my_half == 0) + group_copy_seconds(my_half == 1)
        turns_gbs = world * 5 * h2d / t_turns / 1e9
        ingest["halves_taking_turns_gbs"] = round(turns_gbs, 1)
        if turns_gbs > 1.05 * all_gbs:
            dist.barrier()  # rank 0 creates the shared words before anybody attaches
            host_bar = ShmBarrier(rank, world)

            def e2e_steps_arbitrated(n):
                tickets = []
                for k in range(n):
                    for half in (0, 1):
                        host_bar.wait()
                        if half == my_half:
                            tickets.append(rx.submit(h_iq[k % 2].data_ptr(), B, hsets[k % 2]["out"]))
                            rx.wait_ingest(tickets[-1])
                            if k >= 1:
                                rx.wait(tickets[k - 1])
                rx.wait(tickets[-1])

            rx.reset()
            e2e_steps_arbitrated(max(2, args.warmup))
            barrier()
            t0 = time.perf_counter()
            e2e_steps_arbitrated(args.steps)
            torch.cuda.synchronize()
            s_arb = max_over_ranks(time.perf_counter() - t0)
            barrier()
            host_bar.close()
            arb_value = units / s_arb / 1e6
            ingest["arbitrated_msps"] = round(arb_value, 1)
            if arb_value > e2e_value:
                ingest["mode"] = "arbitrated: the even and the odd ranks take turns on the host link (fmrx_batch_wait_ingest + a shared-memory host barrier)"
                e2e_value, s_e2e = arb_value, s_arb
    ach_gbs = world * S * B * BLOCK_BYTES * args.steps / s_e2e / 1e9
    ceiling = max(all_gbs, ingest.get("halves_taking_turns_gbs", 0.0)) if ingest["mode"] != "free-running" else all_gbs
    link = {"h2d_copy_alone_gbs": round(link_gbs, 1), "all_ranks_copying_gbs": round(all_gbs, 1), "h2d_achieved_gbs": round(ach_gbs, 1),
            "frac_of_link": round(ach_gbs / all_gbs, 3), "frac_of_ceiling_in_use": round(ach_gbs / ceiling, 3), "e2e_ceiling_msps": round(ceiling / 2.0 * 1e3, 1),
            "note": "whole job; h2d_copy_alone is this rank's buffer copied back to back while the other ranks do the same without a barrier, all_ranks_copying the "
                    "same between barriers (the box's ceiling for this many GPUs ingesting at once); the end-to-end path moves 2 bytes per complex sample "
                    "over PCIe, so that rate / 2 is its ceiling; frac_of_link = achieved / all_ranks_copying (above 1 when the ranks take turns on the link, "
                    "config.e2e_ingest); frac_of_ceiling_in_use = achieved / the copy rate of the ingest mode in use"}

    # ---- the same stations through other configurations of the chain, device-resident (each its own handle, same input bytes)
    def side_figure(n_streams, steps, **kw):
        with fmrx.Batch(n_streams, mode=0, max_blocks=B, device=local_rank, **kw) as rxs:
            for _ in range(max(3, args.warmup)):
                rxs.process_device(d_iq.data_ptr(), B, None)
            rxs.sync()
            barrier()
            s_first = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream_phase(rxs.h, 0), device=dev)
            s_last = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream(rxs.h), device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_first)
            for _ in range(steps):
                rxs.process_device(d_iq.data_ptr(), B, None)
            e1.record(s_last)
            rxs.sync()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            return ms / steps, dict(zip(("pll_sms", "filter_sms"), rxs.partition()))

    side_steps = max(10, args.steps // 2)
    sides = {}
    if not args.skip_mode1:
        # the north_star's own target figure: the batched MONO front end (front end + mono low-pass + quantiser).  `binary` profile with
        # the audio path only: the stereo branch is dead after block 0 there, so every timed step is mono only
        ms, part = side_figure(S, side_steps, profile=fmrx.PROFILE_BINARY, paths=fmrx.PATH_AUDIO)
        v = world * S * B * BLOCK_IQ / (ms * 1e-3) / 1e6
        sides["mono_only"] = {"value_msps": round(v, 1), "ms_per_step": round(ms, 4), "realtime_streams": int(v / 2.4), "paths": "front end + mono low-pass + quantiser (binary profile, audio path only)",
                              "target": "north_star: >= 1000 Msps per B200 on the batched mono front end", "sm_partition": part}
        for name, num in (("numerics_fma", fmrx.NUMERICS_FMA), ("numerics_strict", fmrx.NUMERICS_STRICT), ("numerics_reference", fmrx.NUMERICS_REFERENCE)):
            if num == numerics:
                continue
            ms, part = side_figure(S, side_steps, profile=fmrx.PROFILE_INTENT, numerics=num)
            v = world * S * B * BLOCK_IQ / (ms * 1e-3) / 1e6
            sides[name] = {"value_msps": round(v, 1), "ms_per_step": round(ms, 4), "sm_partition": part}
        sides["numerics_note"] = ("headline numerics = %s.  reference: audio path and both band-pass filters ahead of / beside the PLLs with the reference's two roundings per tap; "
                                  "strict: also pllCombine's filter with its double products (whole RDS branch up to the mixer product bit-identical); fma: fused multiply-add "
                                  "wherever the 1e-5 tolerance allows (include/fmrx.h)" % args.numerics)
    # ---- BASELINE config 5 as written: 4096 stations in TOTAL, dealt over the ranks (strong scaling); device-resident and end to end
    strong = None
    if world > 1 and not args.skip_mode1:
        mine = shard.shard_range(args.stations, world, rank)
        Sn = len(mine)
        ms, part = side_figure(Sn, side_steps, profile=fmrx.PROFILE_INTENT, numerics=numerics)
        v = args.stations * B * BLOCK_IQ / (ms * 1e-3) / 1e6
        strong = {"stations_total": args.stations, "stations_per_gpu": Sn, "value_msps": round(v, 1), "ms_per_step": round(ms, 4), "realtime_streams": int(v / 2.4),
                  "sm_partition": part, "scaling": "strong", "note": "per-GPU batches below 1024 stations sit on the PLL's latency floor (one 64 ms block = 15360 serial steps)"}
        with fmrx.Batch(Sn, mode=0, profile=fmrx.PROFILE_INTENT, max_blocks=B, device=local_rank, numerics=numerics) as rxs:
            def steps_e2e(n):
                tk = []
                for k in range(n):
                    tk.append(rxs.submit(h_iq[k % 2].data_ptr(), B, hsets[k % 2]["out"]))
                    if k >= 1:
                        rxs.wait(tk[k - 1])
                rxs.wait(tk[-1])
            steps_e2e(3)
            barrier()
            t0 = time.perf_counter()
            steps_e2e(side_steps)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
        strong["e2e_msps"] = round(args.stations * B * BLOCK_IQ * side_steps / dt / 1e6, 1)

    # ---- the other mode (the metric is per mode): mode 1 = 2.5 Msps, x24 / 125 polyphase audio resamplers, mono + stereo, no RDS
    # (src/fm_radio.cpp:174-180,324); device-resident, same batch size, reported beside the headline in config.mode1
    def other_mode(mode, rate, paths):
        d_iqm = synth.synth_batch_torch(stations, B, mode, dev, chunk=64)
        with fmrx.Batch(S, mode=mode, profile=fmrx.PROFILE_INTENT, max_blocks=B, device=local_rank) as rxm:
            checked = "skipped"
            if rank == 0 and not args.no_check:  # the same spot check as the headline mode: first block(s) of a few stations against the oracle
                from oracle import Chain

                d_a = torch.zeros((S, B, 2 * rxm.n_audio), dtype=torch.int16, device=dev)
                rxm.process_device(d_iqm.data_ptr(), B, fmrx.Outputs(ptr(d_a, fmrx.i16p), None, None, None, None, None))
                rxm.sync()
                for s in sorted({0, min(63, S - 1), S - 1}):
                    ref_audio = Chain(mode, 1).run(d_iqm[s].cpu().numpy())[0]
                    assert np.array_equal(d_a[s].cpu().numpy().ravel(), ref_audio), f"mode {mode} station {s}: audio differs from the oracle"
                checked = "audio bit-exact vs oracle on stations {0,63,S-1}"
                del d_a
                rxm.reset()
            for _ in range(args.warmup):
                rxm.process_device(d_iqm.data_ptr(), B, None)
            rxm.sync()
            barrier()
            s_first = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream_phase(rxm.h, 0), device=dev)
            s_last = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream(rxm.h), device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_first)
            for _ in range(args.steps):
                rxm.process_device(d_iqm.data_ptr(), B, None)
            e1.record(s_last)
            rxm.sync()
            barrier()
            msm = max_over_ranks(e0.elapsed_time(e1))
            vm = units / (msm * 1e-3) / 1e6
            out = {"value_msps": round(vm, 1), "ms_per_step": round(msm / args.steps, 4), "realtime_streams": int(vm / rate), "rf_msps_per_stream": rate, "paths": paths, "parity_spot_check": checked,
                   "sm_partition": dict(zip(("pll_sms", "filter_sms"), rxm.partition()))}
        del d_iqm
        torch.cuda.empty_cache()
        return out

    # ---- one station alone (BASELINE configs 1-4 are single-stream): host bytes in, host audio + RDS out, one 64 ms block
    # per call through the synchronous entry point -- the latency a live receiver sees per block
    single = None
    if rank == 0 and not args.skip_mode1:
        one = d_iq[0].cpu().numpy()
        with fmrx.Batch(1, mode=0, profile=fmrx.PROFILE_INTENT, max_blocks=1, device=local_rank) as r1:
            for _ in range(3):
                r1.process(one[:BLOCK_BYTES])
            t0 = time.perf_counter()
            for _ in range(20):
                r1.process(one[:BLOCK_BYTES])
            ms_blk = (time.perf_counter() - t0) / 20 * 1e3
        single = {"ms_per_block": round(ms_blk, 3), "block_ms_of_signal": 64.0, "realtime_factor": round(64.0 / ms_blk, 1),
                  "msps": round(BLOCK_IQ / (ms_blk * 1e-3) / 1e6, 1), "path": "fmrx_batch_process, 1 station x 1 block per call, host buffers, mono+stereo+rds"}

    mode1 = mode2 = None
    if not args.skip_mode1:
        rx.close()
        del d_iq, h_iq, hsets
        torch.cuda.empty_cache()
        mode1 = other_mode(1, 2.5, "mono+stereo (x24/125 resamplers), 48 kHz")
        mode2 = other_mode(2, 2.4, "mono+stereo (x147/800 resamplers) + rds, 44.1 kHz (not in the reference's main(); BASELINE config 2)")

    if rank != 0:
        return
    # ---- rooflines, every denominator measured in this run (MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only)
    peak_ffma = fmrx.measure_fp32_peak(0, local_rank)      # T FMA/s  -> 2 flop each
    peak_muladd = fmrx.measure_fp32_peak(1, local_rank)    # T lane-ops/s, 1 flop each (the reference-exact tap)
    rate_dfma = fmrx.measure_fp32_peak(4, local_rank)      # T DFMA/s, multiplier and addend uniform (the form the microbenchmark of round 1 measured)
    rate_dfma3 = fmrx.measure_fp32_peak(8, local_rank)     # T DFMA/s with three vector-register operands: what the PLL step's polynomials issue
    rate_mixed = fmrx.measure_fp32_peak(7, local_rank)     # 4 DFMA + 2 conversions interleaved: the two pipes overlap (rate ~ the slower of the two alone)
    rate_cvt = fmrx.measure_fp32_peak(5, local_rank)       # T float<->double conversions/s
    rate_alu = fmrx.measure_fp32_peak(6, local_rank)       # T integer ALU lane-ops/s
    hbm_peak, hbm_src = measured_peaks()
    prof, prof_path = ncu_profile()
    mixes, mix_path = sass_mix()
    fir_order = [n for n in ("mono", "pilot_bpf", "stereo_bpf", "rds_bpf", "rds_sq_bpf", "stereo_lpf") if stage.get(n, (0, 0))[1]]
    exact_now = {"rds_bpf": numerics != fmrx.NUMERICS_FMA, "rds_sq_bpf": False, "stereo_bpf": numerics != fmrx.NUMERICS_FMA}
    # band-pass filters of the discriminator output that share one launch: the stereo and RDS band under FMA numerics (the pilot filter
    # stays exact and on its own); under REFERENCE / STRICT the three exact filters only with FMRX_BPF_FUSED=1 (measured slower)
    n_fused = 2 if numerics == fmrx.NUMERICS_FMA else 3
    STAGE_MACS["bpf_fused"] = (n_fused * NIF * NT, numerics != fmrx.NUMERICS_FMA)
    STAGE_BYTES["bpf_fused"] = 4 * NIF + n_fused * 4 * NIF
    per = {}
    for name, (ms, cnt) in stage.items():
        if cnt == 0:
            continue
        ent = {"ms_per_step": round(ms / args.steps, 4), "share": round(ms / sum(v[0] for v in stage.values()), 4)}
        if name in STAGE_MACS and not (name == "rds_sq_bpf" and numerics == fmrx.NUMERICS_STRICT):
            macs, exact = STAGE_MACS[name]
            exact = exact_now.get(name, exact)
            tf = 2.0 * macs * S * B * args.steps / (ms * 1e-3) / 1e12
            pk = peak_muladd if exact else 2.0 * peak_ffma
            ent.update({"tflops": round(tf, 2), "fp32_peak_tflops": round(pk, 2), "frac_fp32": round(tf / pk, 4), "rounding": "mul+add (reference-exact)" if exact else "fma"})
        if name == "rds_sq_bpf" and numerics == fmrx.NUMERICS_STRICT:
            # double products rounded into a float sum: 4 FP64 instructions per tap (DMUL, DADD and the DADD pair that rounds to float)
            taps = NIF * NT * S * B * args.steps
            ent.update({"fp64_ops_per_tap": 4, "tera_fp64_ops": round(4 * taps / (ms * 1e-3) / 1e12, 2), "fp64_peak_tera_ops": round(rate_dfma, 2),
                        "frac_fp64": round(4 * taps / (ms * 1e-3) / 1e12 / rate_dfma, 4), "rounding": "double product into a float sum (src/helper.cpp:139)"})
        if name == "rds_symbols":
            # SURVEY 8(d) counts the three stages this kernel pair replaces (mixer LPF 15360 x 151, resampler and RRC 3648 x 151
            # each = 22.3 MAC per complex sample); against that figure -- the work the reference does for the same symbols --
            ref_macs = (NIF + 2 * NRDS) * NT
            tf_ref = 2.0 * ref_macs * S * B * args.steps / (ms * 1e-3) / 1e12
            ent.update({"tflops_survey_algorithmic": round(tf_ref, 2), "frac_fp32_survey_algorithmic": round(tf_ref / (2.0 * peak_ffma), 4),
                        "note": "tflops / frac_fp32 count the MACs this formulation executes (composite filter at the decoder's sampling instants); "
                                "the *_survey_algorithmic pair counts the MACs of the three full-rate stages it replaces"})
        gbs = STAGE_BYTES[name] * S * B * args.steps / (ms * 1e-3) / 1e9
        tr = stage_traffic(prof, name, fir_order) if args.numerics == "reference" else None  # the committed capture is of the default numerics
        ent["ncu_pipes_pct"] = stage_pipes(prof, name, fir_order) if args.numerics == "reference" else None
        ent.update({"hbm_gbs": round(gbs, 1), "frac_hbm": round(gbs / hbm_peak, 4), "algorithmic_bytes": STAGE_BYTES[name] * S * B,
                    "traffic": None if tr is None else round(tr * S * B / 4096.0), "traffic_over_algorithmic": None if tr is None else round(tr / (STAGE_BYTES[name] * 4096.0), 3)})
        per[name] = ent
    # ---- the dominant kernel: pll_kernel.  It is not a throughput kernel on the whole device -- 8192 loops = 256 warps on 592
    # schedulers, one dependency chain per sample -- so it runs on its own SM partition (2 warps per scheduler on 32 SMs) beside the
    # filters, and what bounds it THERE is instruction issue: one step is `total` warp-instructions of which `fp64` go down the FP64
    # pipe and `cvt` + `mufu` down the XU pipe (static SASS count, profiles/*_sass_pll_kernel.txt).  The bound per warp-step on a
    # scheduler is the largest of the per-pipe issue times at the rates measured above and the one-instruction-per-cycle issue slot.
    pll = None
    if "pll" in per and "pll_kernel" in mixes:
        mx, spi = mixes["pll_kernel"]["per_iteration"], mixes["pll_kernel"]["steps_per_iteration"]
        sms_total = torch.cuda.get_device_properties(local_rank).multi_processor_count
        clk = peak_ffma * 1e12 / (sms_total * 128.0)                           # effective SM clock during the microbenchmarks, Hz (FFMA: 128 lanes per clock per SM)
        per_sched = lambda rate: rate * 1e12 / (sms_total * 4.0) / clk / 32.0  # noqa: E731  warp-instructions per cycle per scheduler on that pipe
        cyc = {"issue": mx["total"] / spi,
               "fp64": mx.get("fp64", 0) / spi / per_sched(rate_dfma3),
               "xu": (mx.get("cvt", 0) + mx.get("mufu", 0)) / spi / per_sched(rate_cvt),
               "alu": mx.get("alu", 0) / spi / per_sched(rate_alu),
               "fma": (mx.get("fp32", 0) + mx.get("fp32_packed", 0)) / spi / per_sched(peak_ffma)}
        bound_cyc = max(cyc.values())
        lanes = S * 2
        warps = (lanes + 31) // 32
        steps_total = NIF * B
        use_sms = pll_sms if pll_sms else sms_total
        t_used = pll_pipelined_ms if (pll_pipelined_ms and pll_sms) else per["pll"]["ms_per_step"]
        # time the partition's schedulers need for all warp-steps at the bound; a scheduler cannot be shared below one warp
        scheds = use_sms * 4
        t_bound_ms = -(-warps // scheds) * steps_total * bound_cyc / clk * 1e3  # ceil(warps / schedulers) warps take turns on the busiest scheduler
        ach = lanes * steps_total / (t_used * 1e-3) / 1e9
        pk = lanes * steps_total / (t_bound_ms * 1e-3) / 1e9
        chain = fmrx.measure_pll_chain(local_rank)
        pll = {"kernel": "pll_kernel", "bound": "issue", "achieved": round(ach, 2), "peak": round(pk, 2), "unit": "G loop-steps/s", "frac": round(ach / pk, 4),
               "ms_per_launch": round(t_used, 4), "measured": "CUDA events around the kernel on its stream, inside the pipelined run (timeline pass, %d steps), on its %d-SM partition" % (n_tl, use_sms) if (pll_pipelined_ms and pll_sms)
               else "serialised stage pass on the whole device",
               "sms": use_sms, "warps": warps, "warps_per_scheduler": round(warps / scheds, 2), "steps_per_launch": steps_total,
               "scheduler_cycles_per_warp_step": round(t_used * 1e-3 * clk / steps_total / max(1, -(-warps // scheds)), 1),
               "bound_cycles_per_warp_step": round(bound_cyc, 1), "bound_by_pipe": {k: round(v, 1) for k, v in cyc.items()},
               "instructions_per_step": {k: round(v / spi, 2) for k, v in mx.items()}, "instruction_mix_source": mix_path,
               "pipe_rates_tera_lane_ops": {"ffma": round(peak_ffma, 2), "dfma_uniform_operands": round(rate_dfma, 2), "dfma_three_registers": round(rate_dfma3, 2), "f2f": round(rate_cvt, 2),
                                            "alu": round(rate_alu, 2), "dfma_and_f2f_interleaved_4_to_2": round(rate_mixed, 2)},
               "clock_hz_effective": round(clk), "alone_on_whole_device_ms": per["pll"]["ms_per_step"],
               "chain_latency_cycles": round(chain, 1), "latency_floor_ms": round(steps_total * chain / clk * 1e3, 3),
               "note": "alone on the whole device the kernel sits on its latency floor (one warp per scheduler at most: alone_on_whole_device_ms vs latency_floor_ms); "
                       "on the partition two warps share a scheduler and the issue bound is the one that matters",
               "traffic": per["pll"]["traffic"], "algorithmic_bytes": per["pll"]["algorithmic_bytes"], "traffic_source": prof_path}
        per["pll"].update({"cycles_per_step_alone": round(per["pll"]["ms_per_step"] * 1e-3 * clk / steps_total, 1), "chain_latency_cycles": round(chain, 1)})
    top = max((n for n in per if n in STAGE_MACS and "tflops" in per[n]), key=lambda n: per[n]["ms_per_step"])
    fir_roof = {
        "bound": "fp32", "kernel": top, "achieved": per[top]["tflops"], "peak": per[top]["fp32_peak_tflops"], "unit": "TFLOP/s", "frac": per[top]["frac_fp32"],
        "traffic": per[top]["traffic"], "traffic_unit": "bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu capture, " + prof_path,
        "algorithmic_bytes": per[top]["algorithmic_bytes"],
        "peak_source": "measured in this run by fmrx_measure_fp32_peak: %.2f T FFMA/s (x2 flop), %.2f T FMUL+FADD lane-ops/s; a stage that keeps the "
                       "reference's two roundings per tap is bounded by the latter" % (peak_ffma, peak_muladd),
        "hbm": {"achieved": per[top]["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": per[top]["frac_hbm"], "peak_source": hbm_src},
    }
    if pll:
        roofline = dict(pll)
        roofline.update({"largest_filter_kernel": fir_roof, "stages": per})
    else:
        roofline = dict(fir_roof)
        roofline["stages"] = per
    chain_alg = sum(STAGE_BYTES[n] for n in per) * S * B
    chain_tr = [per[n]["traffic"] for n in per]
    roofline["chain_traffic"] = {"algorithmic_bytes_per_step": chain_alg, "traffic_bytes_per_step": None if any(t is None for t in chain_tr) else int(sum(chain_tr)),
                                 "fused_floor_bytes_per_step": int(2.08 * BLOCK_IQ * S * B), "source": prof_path}
    cores = host_cores()
    cpu_msps, cpu_kind, cpu_dt = cpu_reference_run(args.cpu_blocks, cores)
    cpu1_msps, _, cpu1_dt = cpu_reference_run(args.cpu_blocks, 1)
    try:
        py_models = python_models_baseline()
    except Exception as e:  # same rule as the per-stage table below
        py_models = {"error": repr(e)}
    try:
        cpu_stages, cpu_stage_kind = cpu_stage_times()
    except Exception as e:  # the per-stage table is an explanation, not a result: never let it take the bench line down
        cpu_stages, cpu_stage_kind = {"error": repr(e)}, "unavailable"
    line = {
        "metric": "IQ Msps (complex samples/s, whole job)", "value": round(value, 1), "unit": "Msps", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batch4096_mode0_stereo_rds", "stations_per_gpu": S, "blocks_per_step": B, "mode": 0, "profile": "intent", "paths": "mono+stereo+rds",
                   "numerics": args.numerics, "mono_only": sides.get("mono_only"), "numerics_fma": sides.get("numerics_fma"), "numerics_strict": sides.get("numerics_strict"),
                   "numerics_reference": sides.get("numerics_reference"), "numerics_note": sides.get("numerics_note"), "strong_4096_total": strong,
                   "sm_partition": {"pll_sms": pll_sms, "filter_sms": filter_sms} if pll_sms else "none (phases share the device)", "realtime_streams": int(value / 2.4), "e2e_realtime_streams": int(e2e_value / 2.4),
                   "l2": "input per step %.2f GB >> 126 MB L2, no flush needed" % (S * B * BLOCK_BYTES / 1e9), "input_reuse": "same synthesised block replayed each step, state carried",
                   "synth_seconds": round(t_synth, 2), "parity_spot_check": parity, "rank0_numa_node": numa, "gpu_map": gpu_map,
                   "e2e_timer": "host clock around K fmrx_batch_submit calls with fmrx_batch_wait on the previous step (two steps in flight), barrier + synchronize on both sides, max over ranks",
                   "e2e_sync_call_msps": round(e2e_sync_value, 1), "e2e_link": link, "e2e_ingest": ingest, "single_stream": single, "mode1": mode1, "mode2_44k1": mode2},
        "e2e": {"value": round(e2e_value, 1), "unit": "Msps", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world},  # whole job, like `value`
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": {"value": round(cpu_msps, 3), "unit": "Msps", "cores": cores, "kind": cpu_kind,
                         "sample": f"{cores} concurrent processes x {args.cpu_blocks} blocks of the same mode-0 workload ({cpu_dt:.1f} s wall)",
                         "single_process": {"value": round(cpu1_msps, 3), "unit": "Msps", "sample": f"1 process x {args.cpu_blocks} blocks ({cpu1_dt:.1f} s)"},
                         "python_models": py_models,
                         "stages_ms_per_block_one_core": cpu_stages, "stages_kind": cpu_stage_kind,
                         "stages_note": "the reference's functions on one host core, one mode-0 block per call (SURVEY 8d ii); hold against "
                                        "roofline.stages[*].ms_per_step / stations_per_gpu for the GPU's time per station-block"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100, help="timed steps; the default follows SURVEY 8(d): throughput sustained over >= 100 blocks per stream "
                    "(the three-phase pipeline needs about two steps to fill, which a 10-step run still shows as +8 %% per step)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fmrx", choices=["fmrx", "reference"])
    ap.add_argument("--stations", type=int, default=4096, help="stations per GPU")
    ap.add_argument("--blocks", type=int, default=1, help="blocks per station per step")
    ap.add_argument("--cpu-blocks", type=int, default=48, help="blocks per process of the CPU baseline sample")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--numerics", default="reference", choices=["reference", "strict", "fma"], help="include/fmrx.h FMRX_NUMERICS_*")
    ap.add_argument("--skip-mode1", action="store_true", help="do not measure the secondary figures of the other modes (config.mode1, config.mode2_44k1)")
    ap.add_argument("--skip-e2e", action="store_true", help="for the ncu launch list: stop after the device-resident and per-stage passes (ncu's "
                    "measurement library fails with LaunchFailed on the first kernel that waits on an event recorded in another context's stream, "
                    "which is how the asynchronous host path chains its H2D copy to the partitioned filter stream)")
    ap.add_argument("--device-only", action="store_true", help="tuning aid: device-resident number and stage times only (not a bench line)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_fmrx_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
