#!/bin/bash
# One GPU-box pass: parity suite, bench line, ncu launch list of the bench command, one `--set full` capture of every
# kernel of one chain step (4096 stations x 1 block).  Usage (under gpurun): bash tools/gpu_profile.sh <tag>
# Outputs land in gpurun_out/<tag>_*; text summaries are made on the box too so that they survive a dropped report.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?" | tee -a $OUT/${TAG}_tests.log
python bench.py > $OUT/${TAG}_bench.log 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
tail -c 600 $OUT/${TAG}_bench.log
# launch list of the bench command (cold-cache, serialised: compare shares)
python bench.py --steps 2 --warmup 1 --no-check --skip-e2e > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:fmrx --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-check --skip-e2e > $OUT/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
# full capture of one chain step
python tools/prof_chain.py 4096 1 > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fmrx -c 40 \
    -f -o $OUT/${TAG}_chain python tools/prof_chain.py 4096 1 > $OUT/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la $OUT/${TAG}_chain.ncu-rep
ncu -i $OUT/${TAG}_chain.ncu-rep --page raw --csv > $OUT/${TAG}_chain_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_chain.ncu-rep --page details > $OUT/${TAG}_chain_details.txt 2>/dev/null
# keep what comes back under the 64 MiB cap
if [ $(stat -c %s $OUT/${TAG}_chain.ncu-rep) -gt 50000000 ]; then rm -f $OUT/${TAG}_chain.ncu-rep; echo "report too large, dropped (text kept)"; fi
