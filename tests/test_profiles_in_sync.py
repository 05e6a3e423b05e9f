"""The tracked evidence bench.py computes its rooflines from must describe the code that is built: the static instruction mix of the hot
loops (profiles/<tag>_sass_mix.json, tools/sass_mix.py) is recomputed from the objects in the build tree and compared with the committed
file; the tracked ncu summary must hold a row for every kernel of a chain step that the bench maps a stage to."""
import csv
import importlib.util
import json
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump (CUDA toolkit) not on PATH")
def test_committed_instruction_mix_is_the_built_kernels():
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_for_test")
    sm = _load(os.path.join(ROOT, "tools", "sass_mix.py"), "sass_mix_for_test")
    committed, path = bench.sass_mix()
    assert committed, f"{path} is missing"
    obj = os.path.join(sm.BUILD, "fmrx_pll.o")
    if not os.path.exists(obj):
        pytest.skip("build tree absent (objects are not shipped): run __graft_entry__.build()")
    import re
    fns = sm.functions(obj)
    name = next(n for n in fns if re.search(r"pll_kernelILi2", n))
    _, _, _, keep = sm.hot_loop(fns[name], "fp64")
    assert sm.mix(keep) == committed["pll_kernel"]["per_iteration"], "profiles/*_sass_mix.json is stale: python tools/sass_mix.py <tag>"


def test_tracked_ncu_summary_covers_every_bench_stage():
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_for_test2")
    prof, path = bench.ncu_profile()
    assert prof, f"{path} is missing"
    order = ["mono", "pilot_bpf", "stereo_bpf", "rds_bpf", "rds_sq_bpf", "stereo_lpf"]
    for stage in ["frontend", "pll", "combine", "rds_decode", "rds_symbols"] + order:
        assert bench.stage_traffic(prof, stage, order), f"no launch of stage {stage} in {path}"
        assert bench.stage_pipes(prof, stage, order)["busiest"], stage
    rows = list(csv.DictReader(open(os.path.join(ROOT, path))))
    assert all(float(r["duration_ms"]) > 0 for r in rows)
