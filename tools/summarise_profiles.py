"""Turns what a GPU-box pass left in gpurun_out/ into the small text files kept under profiles/.

    python tools/summarise_profiles.py <tag> [--note "..."]

  gpurun_out/<tag>_chain_raw.csv or <tag>_raw.csv  (ncu --page raw --csv of a `--set full` capture)  -> profiles/<tag>_kernels.md
  gpurun_out/<tag>_launches.csv                    (ncu --metrics gpu__time_duration.sum launch list) -> profiles/<tag>_launches.txt
  gpurun_out/<tag>_bench.log                       (the bench line of the same pass)                  -> profiles/<tag>_bench.json
"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("sm__icc_request_hit_rate.pct", "icache hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_inst"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
]


def short(name):
    name = name.replace("void ", "").replace("fmrx::<unnamed>::", "").replace("unnamed>::", "").replace("(int)", "").replace("(bool)", "")
    return name.split("(")[0]


def kernels(tag, note):
    src = next((p for p in (os.path.join(G, f"{tag}_chain_raw.csv"), os.path.join(G, f"{tag}_raw.csv")) if os.path.exists(p)), None)
    if not src:
        return
    rows = list(csv.reader(open(src)))
    h, units = rows[0], rows[1]
    idx = [(h.index(c), lab) for c, lab in COLS if c in h]
    out = [f"# ncu `--set full --clock-control none` capture `{tag}` — one steady-state chain step, 4096 stations x 1 block of the synthetic multiplex (tools/gpu_final.sh: third step of tools/prof_chain.py 4096 3)", ""]
    if note:
        out += [note, ""]
    out += ["Times are ncu's serialised, cold-cache per-launch durations (compare shares, not absolutes); DRAM bytes are per launch.", ""]
    out.append("| kernel | " + " | ".join(f"{lab} [{units[i]}]" if units[i] else lab for i, lab in idx) + " |")
    out.append("|---|" + "---|" * len(idx))
    for r in rows[2:]:
        out.append("| `" + short(r[h.index("Kernel Name")]) + "` | " + " | ".join(r[i] for i, _ in idx) + " |")
    open(os.path.join(P, f"{tag}_kernels.md"), "w").write("\n".join(out) + "\n")
    print("wrote", f"profiles/{tag}_kernels.md", len(rows) - 2, "kernels")


def launches(tag):
    src = os.path.join(G, f"{tag}_launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit() and r[-1].replace(".", "", 1).isdigit()]
    agg = OrderedDict()
    for r in rows:
        k = short(r[4])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + float(r[-1]))
    tot = sum(t for _, t in agg.values())
    out = [f"ncu --metrics gpu__time_duration.sum --clock-control none launch list `{tag}`: {len(rows)} launches of this library's kernels, {tot / 1e6:.3f} ms in total",
           "(cold-cache, serialised: the SHARES are what is comparable with the CUDA-event stage times in the bench line)", "",
           f"{'kernel':72s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k[:72]:72s} {n:8d} {t / 1e3:12.1f} {t / n / 1e3:10.1f} {100 * t / tot:6.1f}%")
    open(os.path.join(P, f"{tag}_launches.txt"), "w").write("\n".join(out) + "\n")
    print("wrote", f"profiles/{tag}_launches.txt")


def bench(tag):
    src = os.path.join(G, f"{tag}_bench.log")
    if not os.path.exists(src):
        return
    for line in open(src):
        if line.startswith("{"):
            json.dump(json.loads(line), open(os.path.join(P, f"{tag}_bench.json"), "w"), indent=1)
            print("wrote", f"profiles/{tag}_bench.json")


if __name__ == "__main__":
    tag = sys.argv[1]
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    os.makedirs(P, exist_ok=True)
    kernels(tag, note)
    launches(tag)
    bench(tag)
