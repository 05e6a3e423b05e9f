"""What bounds the end-to-end figure when several GPUs ingest at once: host -> device copies out of page-locked memory, one
GPU alone, every pair together, all together, with default and write-combined host allocations; plus what the box says
about its topology.  One process, one stream per GPU; device-timed (events on each GPU's stream, wall clock around the
concurrent groups).

    python tools/h2d_probe.py [MiB per copy] > gpurun_out/<tag>_h2d_probe.json
"""
import ctypes as C
import itertools
import json
import os
import subprocess
import sys
import time

import torch

MIB = int(sys.argv[1]) if len(sys.argv) > 1 else 512
REPS = 4


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # noqa: BLE001
        return repr(e)


def main():
    n = torch.cuda.device_count()
    out = {"visible_gpus": n, "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES"), "cpus": len(os.sched_getaffinity(0)), "mib_per_copy": MIB}
    out["topo"] = sh("nvidia-smi topo -m")
    out["numa"] = sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'")
    out["gpu_numa"] = {i: sh(f"cat /sys/bus/pci/devices/{torch.cuda.get_device_properties(i).pci_domain_id:04x}:{torch.cuda.get_device_properties(i).pci_bus_id:02x}:{torch.cuda.get_device_properties(i).pci_device_id:02x}.0/numa_node").strip()
                       for i in range(n)}
    nbytes = MIB << 20
    dev, streams, host, host_wc = [], [], [], []
    for i in range(n):
        torch.cuda.set_device(i)
        dev.append(torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}"))
        streams.append(torch.cuda.Stream(device=i))
        host.append(torch.empty(nbytes, dtype=torch.uint8).pin_memory())
        host[-1].fill_(i + 1)
    # write-combined allocations through the runtime (torch has no switch for it)
    wc_ok = True
    try:
        rtlib = None
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                rtlib = C.CDLL(name)
                break
            except OSError:
                continue
        if rtlib is None:
            import glob

            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*")
            rtlib = C.CDLL(cands[0])
        rtlib.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
        for i in range(n):
            p = C.c_void_p()
            e = rtlib.cudaHostAlloc(C.byref(p), nbytes, 0x04 | 0x01)  # cudaHostAllocWriteCombined | cudaHostAllocPortable
            if e != 0:
                wc_ok = False
                break
            host_wc.append(p.value)
        rtlib.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    except Exception as e:  # noqa: BLE001
        wc_ok = False
        out["wc_error"] = repr(e)

    def copy(i, wc):
        with torch.cuda.stream(streams[i]):
            if wc:
                rtlib.cudaMemcpyAsync(C.c_void_p(dev[i].data_ptr()), C.c_void_p(host_wc[i]), nbytes, 1, C.c_void_p(streams[i].cuda_stream))
            else:
                dev[i].copy_(host[i], non_blocking=True)

    def run(group, wc=False):
        for i in group:  # warm
            torch.cuda.set_device(i)
            copy(i, wc)
        for i in group:
            streams[i].synchronize()
        t0 = time.perf_counter()
        for _ in range(REPS):
            for i in group:
                torch.cuda.set_device(i)
                copy(i, wc)
        for i in group:
            streams[i].synchronize()
        dt = time.perf_counter() - t0
        return round(len(group) * REPS * nbytes / dt / 1e9, 1)

    out["solo_gbs"] = [run([i]) for i in range(n)]
    out["all_gbs"] = run(list(range(n)))
    if n >= 2:
        out["pairs_gbs"] = {f"{i},{j}": run([i, j]) for i, j in itertools.combinations(range(n), 2)}
    if n >= 4:
        out["halves_gbs"] = {"low": run(list(range(n // 2))), "high": run(list(range(n // 2, n))), "even": run(list(range(0, n, 2))), "odd": run(list(range(1, n, 2)))}
        if n == 8:
            best = max(itertools.combinations(range(8), 4), key=lambda g: sum(out["pairs_gbs"][f"{a},{b}"] for a, b in itertools.combinations(g, 2)))
            out["best4_by_pairs"] = {"gpus": list(best), "gbs": run(list(best))}
    if wc_ok:
        out["wc_solo_gbs"] = [run([i], True) for i in range(n)]
        out["wc_all_gbs"] = run(list(range(n)), True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
