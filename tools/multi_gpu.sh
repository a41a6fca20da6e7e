#!/bin/bash
# Multi-GPU validation in one call:  gpurun --gpus N --timeout 900 -- 'bash tools/multi_gpu.sh N TAG'
#   bench.py at N GPUs (parity vs oracle and vs one GPU in the line), BASELINE config 4 (16 Mi members, one
#   UserEvent) and config 5 (two 8 Mi WAN pools) sharded over the N GPUs with a one-GPU replay for the digest.
# Box time is charged N-fold: the CPU baseline of the bench line is skipped here (it idles N GPUs).
set -u
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-mg$N}; OUT=gpurun_out/$TAG; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29611 bench.py --gpus $N --steps 4 --warmup 3 --skip-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
timeout 200 $RUN --master-port 29612 tools/c4_event.py --check > $OUT/c4.json 2> $OUT/c4.err
timeout 240 $RUN --master-port 29613 tools/c5_wan.py --check > $OUT/c5.json 2> $OUT/c5.err
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_sharded.py -q > $OUT/pytest.log 2>&1; fi
for f in bench c4 c5; do echo "== $f"; tail -c 1500 $OUT/$f.json; tail -2 $OUT/$f.err; done
