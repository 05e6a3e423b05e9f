#!/bin/bash
# one GPU call: parity suite, the default bench line, the ncu launch list of the same command and one --set full capture of a chain step
TAG=${1:-r4d}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q -rP > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|error" $OUT/${TAG}_tests.log | tail -3
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench.err
python tools/prof_chain.py 4096 1 > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:fmrx" -c 40 -f -o $OUT/${TAG}_chain python tools/prof_chain.py 4096 1 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i $OUT/${TAG}_chain.ncu-rep --page raw --csv > $OUT/${TAG}_chain_raw.csv 2>/dev/null
python tools/ncu_trim.py $OUT/${TAG}_chain_raw.csv $OUT/${TAG}_ncu_kernels.csv
python tools/prof_chain.py 4096 1 strict > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:sq_exact" -c 1 -f -o $OUT/${TAG}_sq python tools/prof_chain.py 4096 1 strict > $OUT/${TAG}_ncu_sq.log 2>&1
ncu -i $OUT/${TAG}_sq.ncu-rep --page raw --csv > $OUT/${TAG}_sq_raw.csv 2>/dev/null
python tools/ncu_trim.py $OUT/${TAG}_sq_raw.csv $OUT/${TAG}_ncu_sq.csv
ncu -i $OUT/${TAG}_sq.ncu-rep --page details > $OUT/${TAG}_sq_details.txt 2>/dev/null
python bench.py --steps 20 --skip-e2e --no-check > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:fmrx" -s 60 -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 20 --skip-e2e --no-check > $OUT/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
rm -f $OUT/${TAG}_chain_raw.csv $OUT/${TAG}_sq_raw.csv
ls -la $OUT/${TAG}_*
