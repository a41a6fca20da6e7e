#!/bin/bash
# End-of-round verification on one B200:  gpurun --timeout 1500 -- 'bash tools/final_check.sh TAG'
# (tools/configs_report.py and tools/parity_at.py are separate: their results do not depend on kernel speed)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log)
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:gs_window_kernel -s 0 -c 1 -o $OUT/win_1m -f python tools/prof_target.py --members 1000000 --ticks 2048 > $OUT/ncu_win_1m.log 2>&1
timeout 300 $NCU -k regex:gs_window_kernel -s 0 -c 1 -o $OUT/win_64m -f python tools/prof_target.py --members 67108864 --ticks 400 > $OUT/ncu_win_64m.log 2>&1
timeout 200 $NCU -k regex:gs_tick -s 24 -c 1 -o $OUT/tick_cascade_1m -f python tools/prof_target.py --members 1000000 --ticks 64 --join --nograph > $OUT/ncu_tick_cascade_1m.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-hbm-point --skip-cpu-baseline --skip-parity > $OUT/bench_under_ncu.log 2>&1
timeout 100 python tools/e2e_breakdown.py > $OUT/e2e_breakdown.log 2>&1
timeout 200 python tools/win_bench.py consul_b200/libgsim.so > $OUT/win_bench.log 2>&1
GSIM_NO_PRISTINE_WINDOWS=1 timeout 200 python tools/win_bench.py consul_b200/libgsim.so > $OUT/win_bench_per_probe_loop.log 2>&1
tail -3 $OUT/pytest.log; head -c 400 $OUT/bench.json; echo; head -c 600 $OUT/bench_reference.json; echo; cat $OUT/e2e_breakdown.log $OUT/win_bench.log $OUT/win_bench_per_probe_loop.log
