#!/bin/bash
# Diagnostic builds of the attention backward next to the product library (selected with TGAN_B200_LIB=...):
#   libtgan_b200_s4.so    TGAN_BWD_SPLIT=4 (16 row warps, 16 columns per thread)
#   libtgan_b200_prof.so  -DTGAN_PROFILE: per-phase clock64 accumulators (tools/bwd_phase_prof.py)
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/../transformer-gan_b200/csrc" && pwd)"
OBJ="$HERE/build"; OUT="$HERE/../tgan_b200"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
OTHERS=$(ls "$OBJ"/*.o | grep -v "relattn_tc\(_.*\)\?\.o$" | grep -v "_prof.o$")
$NVCC $FLAGS -DTGAN_BWD_SPLIT=4 -c "$HERE/relattn_tc.cu" -o "$OBJ/relattn_tc_s4.o" &
$NVCC $FLAGS -DTGAN_PROFILE -c "$HERE/relattn_tc.cu" -o "$OBJ/relattn_tc_prof.o" &
wait
$NVCC -shared -o "$OUT/libtgan_b200_s4.so" $OTHERS "$OBJ/relattn_tc_s4.o" -lcudart
$NVCC -shared -o "$OUT/libtgan_b200_prof.so" $OTHERS "$OBJ/relattn_tc_prof.o" -lcudart
ls -la "$OUT"/*.so
