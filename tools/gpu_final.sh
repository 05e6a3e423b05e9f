#!/bin/bash
# final pass of a round on one GPU: parity suite, smoke, both bench arms (timed), launch list of the bench command
TAG=${1:-r4t}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q -rP > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|rror" $OUT/${TAG}_tests.log | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err ) 2>&1 | grep real
( time python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -c 300 $OUT/${TAG}_bench.err
python bench.py --steps 20 --skip-e2e --no-check > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:fmrx" -s 60 -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 20 --skip-e2e --no-check > $OUT/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
python tools/rds_noise_agreement.py 2048 10 4 > $OUT/${TAG}_rds_noise.txt 2>&1; cat $OUT/${TAG}_rds_noise.txt
