#!/bin/bash
# ncu evidence for the dopri5 kernels at the C3 (D=12, odeint calls of 10 patients) and C1 (D=6, calls of 50) shapes.
# Usage (under gpurun, one GPU): scripts/prof_dopri5.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
C3="python scripts/kbench_dopri5.py --groups 8192 --batch 10 --D 12"
C1="python scripts/kbench_dopri5.py --groups 2048 --batch 50 --D 6"
$C3 > $OUT/c3_plain.json 2> $OUT/c3_plain.err || { echo "C3 plain run failed"; tail -5 $OUT/c3_plain.err; exit 1; }
$C1 > $OUT/c1_plain.json 2> $OUT/c1_plain.err || { echo "C1 plain run failed"; tail -5 $OUT/c1_plain.err; exit 1; }
cat $OUT/c3_plain.json $OUT/c1_plain.json
C3S="python scripts/kbench_dopri5.py --groups 4096 --batch 10 --D 12 --reps 1"
C1S="python scripts/kbench_dopri5.py --groups 1024 --batch 50 --D 6 --reps 1"
for K in dopri5_fwd dopri5_bwd; do
  if [ $K = dopri5_fwd ]; then SKIP=2; else SKIP=1; fi
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o $OUT/c3_$K $C3S > $OUT/ncu_c3_$K.log 2>&1
  python scripts/ncu_summary.py $OUT/c3_$K.ncu-rep > $OUT/c3_$K.summary.txt 2>> $OUT/ncu_c3_$K.log
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o $OUT/c1_$K $C1S > $OUT/ncu_c1_$K.log 2>&1
  python scripts/ncu_summary.py $OUT/c1_$K.ncu-rep > $OUT/c1_$K.summary.txt 2>> $OUT/ncu_c1_$K.log
  if [ "${KEEP_REP:-0}" != "1" ]; then rm -f $OUT/c3_$K.ncu-rep $OUT/c1_$K.ncu-rep; fi
done
ls -la $OUT
