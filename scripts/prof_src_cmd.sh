#!/bin/bash
# source-level ncu capture of one kernel of an arbitrary command: scripts/prof_src_cmd.sh <tag> <kernel-regex> <skip> <command...>
TAG=$1; K=$2; SKIP=$3; shift 3
OUT=gpurun_out/src_$TAG
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o $OUT/k "$@" > $OUT/ncu.log 2>&1
python scripts/ncu_summary.py $OUT/k.ncu-rep > $OUT/summary.txt 2>> $OUT/ncu.log
ncu -i $OUT/k.ncu-rep --page source --csv --print-source sass > $OUT/sass.csv 2>> $OUT/ncu.log
rm -f $OUT/k.ncu-rep
ls -la $OUT
