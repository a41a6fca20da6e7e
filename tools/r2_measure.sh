#!/bin/bash
# One-call measurement of the prepared kernel variants and launch-shape knobs (one B200, ~10 min):
#   1. parity of the semantic variants on the GPU (torch-free quick check, variant .so swapped in)
#   2. us/tick of every variant and knob at 1 M / 16 Mi / 64 Mi members
#   3. ncu --set full of the tick kernel for the default and the combined variant at 1 M and 64 Mi
# Everything lands in gpurun_out/r2/.  Usage: gpurun --timeout 1100 -- 'bash tools/r2_measure.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2
mkdir -p $OUT
LIBS="consul_b200/libgsim.so consul_b200/libgsim_kstat.so consul_b200/libgsim_mailmap.so consul_b200/libgsim_both.so consul_b200/libgsim_earlya.so consul_b200/libgsim_mb2.so consul_b200/libgsim_mb1.so"
for l in $LIBS; do [ -f $l ] || { echo "missing $l: run tools/build_variants.sh first"; exit 2; }; done

# 2. timing first (the numbers that decide the next step)
python tools/variant_bench.py $LIBS > $OUT/variants.log 2>&1
for k in 1 2; do
  echo "== GSIM_CTAS_PER_SM=$k" >> $OUT/knobs.log
  GSIM_CTAS_PER_SM=$k python tools/variant_bench.py consul_b200/libgsim.so >> $OUT/knobs.log 2>&1
done
echo "== GSIM_MULTI_TICK=1 GSIM_CTAS_PER_SM=1" >> $OUT/knobs.log
GSIM_MULTI_TICK=1 GSIM_CTAS_PER_SM=1 python tools/variant_bench.py consul_b200/libgsim.so >> $OUT/knobs.log 2>&1
echo "== GSIM_NO_PDL=1" >> $OUT/knobs.log
GSIM_NO_PDL=1 python tools/variant_bench.py consul_b200/libgsim.so >> $OUT/knobs.log 2>&1
echo "== GSIM_NO_L2_WINDOW=1 (kstat without the persistence window)" >> $OUT/knobs.log
GSIM_NO_L2_WINDOW=1 python tools/variant_bench.py consul_b200/libgsim_kstat.so >> $OUT/knobs.log 2>&1

# 1. parity of the semantic variants (the quick check links libgsim.so by name: swap the file in a scratch dir)
g++ -O1 -std=c++17 -Iinclude tests/facade/gpu_quickcheck.cpp -Lconsul_b200 -lgsim -Loracle -loracle \
    -Wl,-rpath,'$ORIGIN/lib' -Wl,-rpath,$PWD/oracle -o $OUT/quickcheck
mkdir -p $OUT/lib
for l in consul_b200/libgsim_kstat.so consul_b200/libgsim_mailmap.so consul_b200/libgsim_both.so; do
  cp $l $OUT/lib/libgsim.so
  echo "== $(basename $l)" >> $OUT/parity.log
  timeout 120 $OUT/quickcheck >> $OUT/parity.log 2>&1; echo "rc=$?" >> $OUT/parity.log
done
rm -rf $OUT/lib $OUT/quickcheck

# 3. ncu of the dominant kernel (one launch each; --clock-control none as B200_PROFILING.md says)
for cfg in "default consul_b200/libgsim.so" "both consul_b200/libgsim_both.so"; do
  set -- $cfg
  for n in 1000000 67108864; do
    GSIM_LIB=$2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gs_tick -s 20 -c 1 \
        -o $OUT/prof_$1_$n -f python tools/prof_target.py --members $n --ticks 40 --nograph > $OUT/ncu_$1_$n.log 2>&1
  done
done
tail -n 60 $OUT/variants.log
