#!/bin/bash
# End-of-round verification on one B200:  gpurun --timeout 1500 -- 'bash tools/final_check.sh TAG'
set -u
cd "$(dirname "$0")/.."
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log)
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python tools/parity_at.py 8388352 8388608 > $OUT/parity_8m.json 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:gs_window -s 1 -c 1 -o $OUT/win_1m -f python tools/prof_target.py --members 1000000 --ticks 1000 > $OUT/ncu_win_1m.log 2>&1
timeout 300 $NCU -k regex:gs_window -s 0 -c 1 -o $OUT/win_64m -f python tools/prof_target.py --members 67108864 --ticks 400 > $OUT/ncu_win_64m.log 2>&1
timeout 200 $NCU -k regex:gs_tick -s 24 -c 1 -o $OUT/tick_cascade_1m -f python tools/prof_target.py --members 1000000 --ticks 64 --join --nograph > $OUT/ncu_tick_cascade_1m.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-hbm-point --skip-cpu-baseline --skip-parity > $OUT/bench_under_ncu.log 2>&1
timeout 400 python tools/configs_report.py > $OUT/configs.jsonl 2> $OUT/configs.err
tail -3 $OUT/pytest.log; head -c 400 $OUT/bench.json; echo; head -c 600 $OUT/bench_reference.json; echo; cat $OUT/parity_8m.json | head -c 600; echo; cat $OUT/configs.jsonl
