#!/bin/sh
# Diagnostic builds of the product library (NOT shipped): the TMA-staged kernel without its stores / without its loads,
# to see which side of the memory system bounds a workload.  Use with CVGS_B200_LIB=<path> (results are garbage).
set -e
cd "$(dirname "$0")/../cvgpuspeedup_b200"
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -split-compile 8 -Xcompiler -fPIC -shared"
SRCS="csrc/preproc.cu csrc/circular_tensor.cu csrc/selftest.cu csrc/div_const.o"
mkdir -p ../gpurun_out
nvcc $FLAGS -DCVGS_DIAG_SKIP_STORES -o /tmp/libcvgs_diag_nostore.so $SRCS &
nvcc $FLAGS -DCVGS_DIAG_SKIP_LOADS -o /tmp/libcvgs_diag_noload.so $SRCS &
# 128 registers per thread (four CTAs per SM): add  nvcc $FLAGS -DCVGS_MAX_RESIDENT=4 -o diag/lib_r4.so $SRCS  (measured: c3 40.6 against 41.7 us)
# other cache policies of the output stores: add e.g.  nvcc $FLAGS '-DCVGS_ST_F32="st.global.cg.f32"' -o diag/lib_st_cg.so $SRCS
wait
mkdir -p diag && cp /tmp/libcvgs_diag_nostore.so /tmp/libcvgs_diag_noload.so diag/
