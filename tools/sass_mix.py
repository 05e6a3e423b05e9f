"""Static instruction mix of the hot loops, from the SASS of the objects the library is linked from (cuobjdump -sass on
real-time-software-defined-radio_b200/build/*.o, compiled with -lineinfo).  Writes

    profiles/<tag>_sass_mix.json   per kernel: instructions per loop iteration by pipe class, and what one iteration covers
    profiles/<tag>_sass_<kernel>.txt   the loop bodies themselves (instruction text only), the evidence for the counts

bench.py reads the JSON to turn measured pipe rates into the issue bound of the PLL kernel (roofline.pll).

    python tools/sass_mix.py [tag]

A "hot loop" is the backward branch whose body holds the most instructions of the kernel's dominant class; regions of the
body that a forward branch skips AND that contain a CALL (the out-of-line libm redo path of the PLL step) are left out.
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "real-time-software-defined-radio_b200", "build")

CLASSES = (
    ("fp64", r"^(DFMA|DADD|DMUL|DSETP|DMNMX)"),
    ("cvt", r"^(F2F|F2I|I2F|FRND|F2FP|I2I)"),
    ("mufu", r"^MUFU"),
    ("fp32_packed", r"^(FFMA2|FMUL2|FADD2)"),
    ("fp32", r"^(FFMA|FMUL|FADD|FSET|FMNMX|FSEL|FCHK|IMAD|HFMA2|HADD2|HMUL2)"),
    ("mem", r"^(LDG|STG|LDS|STS|LDC|ULDC|LDL|STL|CCTL|PREFETCH|LDGSTS|LDSM|ATOM|RED)"),
    ("ctrl", r"^(BRA|BSSY|BSYNC|CALL|RET|EXIT|WARPSYNC|BAR|NOP|YIELD|DEPBAR|BMOV|NANOSLEEP)"),
)


def classify(op):
    for name, pat in CLASSES:
        if re.match(pat, op):
            return name
    return "alu"  # LOP3, IADD3, SHF, PRMT, SEL, ISETP, MOV, VIADD, LEA, VIMNMX, ... and the uniform-datapath twins


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            ins = m.group(2).strip()
            pred = re.match(r"^(@!?U?P\d+)\s+(.*)$", ins)
            body = pred.group(2) if pred else ins
            out[name].append((int(m.group(1), 16), body.split()[0].split(".")[0], ins))
    return out


def hot_loop(ins, dominant):
    addr = [a for a, _, _ in ins]
    cands = []
    for a, op, text in ins:
        if op != "BRA":
            continue
        m = re.search(r"(0x[0-9a-f]+)\s*$", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        body = [(x, o, t) for x, o, t in ins if tgt <= x <= a]
        score = sum(1 for _, o, _ in body if classify(o) == dominant)
        cands.append((score, a - tgt, tgt, a, body))
    # a loop without a CALL that holds at least half of the best loop's work wins (the PLL kernel's fast run sits inside an outer
    # loop whose other half is the libm redo path, four inlined re-arm evaluations included)
    top = max(c[0] for c in cands)
    free = [c for c in cands if c[0] >= 0.5 * top and not any(o == "CALL" for _, o, _ in c[4])]
    pool = free if free else [c for c in cands if c[0] >= 0.9 * top]
    _, _, lo, hi, body = min(pool, key=lambda c: c[1])  # the innermost of the loops that hold the work
    # cold regions: forward branches inside the body whose skipped range contains a CALL
    cold = []
    for a, op, text in body:
        if op != "BRA":
            continue
        m = re.search(r"(0x[0-9a-f]+)\s*$", text)
        tgt = int(m.group(1), 16) if m else None
        if tgt and a < tgt <= hi and any(o == "CALL" for x, o, _ in body if a < x < tgt):
            cold.append((a, tgt))
    keep = [(x, o, t) for x, o, t in body if not any(c0 < x < c1 for c0, c1 in cold)]
    return lo, hi, cold, keep


def mix(keep):
    d = {}
    for _, op, _ in keep:
        c = classify(op)
        d[c] = d.get(c, 0) + 1
    d["total"] = len(keep)
    return d


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r5"
    out_dir = os.path.join(ROOT, "profiles")
    spec = (
        # (object, function regex, dominant class, label, what one loop iteration covers)
        ("fmrx_pll.o", r"pll_kernelILi2", "fp64", "pll_kernel", {"steps_per_iteration": 4, "note": "one iteration = four PLL steps of one loop per lane (fast path; the libm redo path is excluded) + the float4 load / store of four samples"}),
        ("fmrx_fir.o", r"frontend_stream4_kernelILb1", "fp32_packed", "frontend_stream4_kernel", {"rows_per_iteration": 4, "note": "one iteration = four rows of ten complex samples: 4 x 151 taps on an (I, Q) pair = 1208 exact taps = 2416 FFMA2 less the taps outside 0..150"}),
        ("fmrx_fir.o", r"fir151_sq_exact_kernel", "fp64", "fir151_sq_exact_kernel", {"taps_per_iteration": 64, "note": "steady-state loop, unrolled by 8 samples x 8 outputs"}),
    )
    result = {}
    for obj, pat, dom, label, meta in spec:
        fns = functions(os.path.join(BUILD, obj))
        name = next(n for n in fns if re.search(pat, n))
        lo, hi, cold, keep = hot_loop(fns[name], dom)
        m = mix(keep)
        result[label] = dict(function=name, loop=[hex(lo), hex(hi)], cold_regions=[[hex(a), hex(b)] for a, b in cold], per_iteration=m, **meta)
        with open(os.path.join(out_dir, f"{tag}_sass_{label}.txt"), "w") as f:
            f.write(f"# {name}\n# hot loop {hex(lo)}..{hex(hi)}, cold regions left out: {[(hex(a), hex(b)) for a, b in cold]}\n# per iteration: {m}\n")
            for a, _, t in keep:
                f.write(f"/*{a:04x}*/ {t}\n")
        print(label, m)
    with open(os.path.join(out_dir, f"{tag}_sass_mix.json"), "w") as f:
        json.dump(result, f, indent=1)


if __name__ == "__main__":
    main()
