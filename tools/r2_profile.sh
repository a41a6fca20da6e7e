#!/bin/bash
# ncu evidence of the round-2 kernels on one B200 (gpurun --timeout 900 -- 'bash tools/r2_profile.sh TAG'):
#   gs_window_kernel at 1 M and 64 Mi members (steady state), gs_tick_kernel in the middle of a join
#   cascade at 1 M members, and the launch list of a short bench run.  Outputs: gpurun_out/$TAG/
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:gs_window -s 6 -c 1 -o $OUT/win_1m -f python tools/prof_target.py --members 1000000 --ticks 400 > $OUT/ncu_win_1m.log 2>&1
timeout 300 $NCU -k regex:gs_window -s 3 -c 1 -o $OUT/win_64m -f python tools/prof_target.py --members 67108864 --ticks 120 > $OUT/ncu_win_64m.log 2>&1
timeout 200 $NCU -k regex:gs_tick -s 24 -c 1 -o $OUT/tick_cascade_1m -f python tools/prof_target.py --members 1000000 --ticks 64 --join --nograph > $OUT/ncu_tick_cascade_1m.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-hbm-point --skip-cpu-baseline --skip-parity > $OUT/bench_under_ncu.log 2>&1
ls -la $OUT
