#!/bin/bash
# final pass of a round on one GPU: parity suite, smoke, both bench arms (timed), launch list of the bench command, one `--set full`
# capture of a steady-state chain step (every kernel), RDS agreement on noisy input
TAG=${1:-r5}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q -rP > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|rror" $OUT/${TAG}_tests.log | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --impl reference --steps 20 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err ) 2>&1 | grep real
( time python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -c 300 $OUT/${TAG}_bench.err
python bench.py --steps 20 --skip-e2e --no-check > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:fmrx" -s 60 -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 20 --skip-e2e --no-check > $OUT/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
python tools/prof_chain.py 4096 3 > $OUT/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fmrx -s 27 -c 13 \
    -f -o $OUT/${TAG}_chain python tools/prof_chain.py 4096 3 > $OUT/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ncu -i $OUT/${TAG}_chain.ncu-rep --page raw --csv > $OUT/${TAG}_chain_raw.csv 2>/dev/null
if [ $(stat -c %s $OUT/${TAG}_chain.ncu-rep) -gt 40000000 ]; then rm -f $OUT/${TAG}_chain.ncu-rep; echo "report too large, dropped (csv kept)"; fi
python tools/rds_noise_agreement.py 2048 10 4 > $OUT/${TAG}_rds_noise.txt 2>&1; tail -12 $OUT/${TAG}_rds_noise.txt
