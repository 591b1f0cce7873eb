#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): scripts/profile_r02.sh
#  (1) scripts/profile.sh r02: plain bench run, launch list, --set full summaries of the fixed-grid kernels of the bench step
#  (2) scripts/prof_dopri5.sh r02: --set full summaries of the dopri5 kernels at the C3 / C1 shapes
set -u
scripts/profile.sh r02
scripts/prof_dopri5.sh r02d
