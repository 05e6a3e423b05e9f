#!/bin/bash
# one `--set full` capture of the kernels matching <regex> in a few chain steps: bash tools/gpu_ncu.sh <tag> <regex> [count] [skip] [steps]
TAG=$1; RE=$2; CNT=${3:-8}; SKIP=${4:-0}; STEPS=${5:-1}
OUT=gpurun_out; mkdir -p $OUT
python tools/prof_chain.py 4096 $STEPS > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$RE" -s $SKIP -c $CNT \
    -f -o $OUT/${TAG} python tools/prof_chain.py 4096 $STEPS > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; ls -la $OUT/${TAG}.ncu-rep
ncu -i $OUT/${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}.ncu-rep --page details > $OUT/${TAG}_details.txt 2>/dev/null
if [ $(stat -c %s $OUT/${TAG}.ncu-rep) -gt 45000000 ]; then rm -f $OUT/${TAG}.ncu-rep; echo "report too large, dropped (text kept)"; fi
