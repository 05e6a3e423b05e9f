"""Cuts an `ncu --page raw --csv` export (2000+ columns, tens of MB) down to the columns bench.py and DESIGN.md quote, one row
per profiled launch, values converted to plain units (bytes, milliseconds, percent):

    python tools/ncu_trim.py gpurun_out/<tag>_raw.csv profiles/<tag>_ncu_kernels.csv

bench.py reads the result for `roofline.*.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch) instead of
carrying literals.  Times in this file are ncu's (cold caches, serialised, clocks not locked): use them for shares, not as
the kernels' durations.
"""
import csv
import re
import sys

COLS = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max": "sm_cycles",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3,
         "nsecond": 1e-6}


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|unnamed>::|<unnamed>::|fmrx::|void ", "", name)
    return re.sub(r"\(.*$", "", name).strip()


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]
    keep = [(c, n) for c, n in COLS.items() if c in idx]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel"] + [n for _, n in keep])
        for k, r in enumerate(data):
            if len(r) <= kn:
                continue
            out = [k, short(r[kn])]
            for c, n in keep:
                v, u = r[idx[c]], units[idx[c]]
                try:
                    x = float(v.replace(",", "")) * SCALE.get(u, 1.0)
                    out.append(f"{x:.6g}")
                except ValueError:
                    out.append(v)
            w.writerow(out)
    print(f"{len(data)} launches -> {dst}")


if __name__ == "__main__":
    main()
