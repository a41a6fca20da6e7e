#!/bin/bash
# Builds kernel variants of libgsim next to the default one, for tools/variant_bench.py on a B200:
#   libgsim_kstat.so   -DGS_KSTAT      status replica for peer gathers + L2 persistence window
#   libgsim_mailmap.so -DGS_MAILMAP    1 mailbox bit per member for the scan (4 B word only when raised)
#   libgsim_both.so    -DGS_KSTAT -DGS_MAILMAP
#   libgsim_earlya.so  -DGS_EARLY_A    L1 prefetch of the probing tile's own columns before the scan
#   libgsim_mb2.so     -DGS_MIN_BLOCKS=2, libgsim_mb1.so -DGS_MIN_BLOCKS=1   fewer, fatter CTAs per SM
# usage: tools/build_variants.sh && python tools/variant_bench.py consul_b200/libgsim.so consul_b200/libgsim_kstat.so ...
set -e
cd "$(dirname "$0")/.."
SRC="consul_b200/csrc/gs_cuda.cu consul_b200/csrc/gs_vmm.cu consul_b200/csrc/gs_api.cpp"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off -shared"
nvcc $FLAGS -DGS_KSTAT=1 -o consul_b200/libgsim_kstat.so $SRC
nvcc $FLAGS -DGS_MAILMAP=1 -o consul_b200/libgsim_mailmap.so $SRC
nvcc $FLAGS -DGS_KSTAT=1 -DGS_MAILMAP=1 -o consul_b200/libgsim_both.so $SRC
nvcc $FLAGS -DGS_EARLY_A=1 -o consul_b200/libgsim_earlya.so $SRC
nvcc $FLAGS -DGS_MIN_BLOCKS=2 -o consul_b200/libgsim_mb2.so $SRC
nvcc $FLAGS -DGS_MIN_BLOCKS=1 -o consul_b200/libgsim_mb1.so $SRC
ls -la consul_b200/libgsim*.so
