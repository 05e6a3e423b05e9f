#!/bin/bash
# A/B of a kernel variant selected by an environment variable: parity suite once, then the device-only bench per value.
#   bash tools/gpu_ab.sh <tag> <ENVVAR> <value>...
TAG=$1; VAR=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_tests.log
for v in "$@"; do
  echo "== $VAR=$v"
  env $VAR=$v python bench.py --device-only --no-check 2>&1 | tee -a $OUT/${TAG}_ab.log | python3 -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(d)
    else: print(l)
"
done
