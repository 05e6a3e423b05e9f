#!/bin/bash
# quick GPU iteration: parity suite + device-only bench under a few PLL partition sizes + launch list of one chain step.
#   bash tools/gpu_iter.sh <tag> [sms...]
TAG=${1:-it}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
for sms in "${@:-32}"; do
  FMRX_PLL_SMS=$sms python bench.py --device-only --no-check 2>&1 | tee -a $OUT/${TAG}_sweep.log
done
python tools/prof_chain.py 4096 2 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:fmrx -s 24 --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/prof_chain.py 4096 2 > $OUT/${TAG}_ncu1.log 2>&1
python3 - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/${TAG}_launches.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:70].ljust(70), r[-1])
PY
