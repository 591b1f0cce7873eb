#!/bin/bash
# ncu evidence for bench.py (run under gpurun, one GPU).  Usage: scripts/profile.sh <tag>
# 1) plain run (must exit 0), 2) launch list with per-launch device time, 3) --set full on the four hot kernels
#    (on a 2^17-patient cohort: ncu saves/restores all device memory around each of its ~40 replays).
set -u
TAG=${1:-r02}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --cpu-patients 512"
$CMD > $OUT/plain.json 2> $OUT/plain.err || { echo "plain run failed"; tail -5 $OUT/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
SMALL="python bench.py --steps 1 --warmup 3 --cpu-patients 128 --patients 131072 --no-extras"
$SMALL > $OUT/plain_small.json 2> $OUT/plain_small.err || { echo "small plain run failed"; tail -5 $OUT/plain_small.err; exit 1; }
# gpurun brings back at most 64 MiB: summarise every capture ON THE BOX (scripts/ncu_summary.py reads the report with
# `ncu -i ... --page raw --csv`) and keep only the text unless KEEP_REP=1
for K in fixed_fwd_sse_kernel fixed_bwd_kernel fixed_fwd_kernel fixed_adj_kernel decode_sse; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $OUT/$K $SMALL > $OUT/ncu_$K.log 2>&1
  python scripts/ncu_summary.py $OUT/$K.ncu-rep > $OUT/$K.summary.txt 2>> $OUT/ncu_$K.log
  if [ "${KEEP_REP:-0}" != "1" ]; then rm -f $OUT/$K.ncu-rep; fi
done
ls -la $OUT
