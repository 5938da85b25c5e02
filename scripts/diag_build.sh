#!/bin/bash
# Diagnostic variants of the library (NOT shipped; the NO_B / NO_MMA ones compute garbage by construction): which
# resource binds the stream-K layer-1 products?  Load one with WGAN_B200_LIB=<file name> (see _lib.py).
#   nob: stream A only     nomma: 1 of 4 MMAs per K block     burstN: producer issues K blocks in groups of N
cd "$(dirname "$0")/.." || exit 1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
S=arshamg_scrnaseq_wgan_b200/csrc/engine.cu
O=arshamg_scrnaseq_wgan_b200
build() { name=$1; shift; nvcc $F "$@" -o $O/libwgan_diag_$name.so $S 2>&1 | grep -E "error" ; }
build nob -DWGAN_DIAG_NO_B &
build nomma -DWGAN_DIAG_NO_MMA &
build nob_nomma -DWGAN_DIAG_NO_B -DWGAN_DIAG_NO_MMA &
build burst2 -DWGAN_PAIR_BURST=2 &
wait
build burst3 -DWGAN_PAIR_BURST=3 &
build nob_nomma_burst2 -DWGAN_DIAG_NO_B -DWGAN_DIAG_NO_MMA -DWGAN_PAIR_BURST=2 &
build nob_nomma_burst3 -DWGAN_DIAG_NO_B -DWGAN_DIAG_NO_MMA -DWGAN_PAIR_BURST=3 &
build nob_burst2 -DWGAN_DIAG_NO_B -DWGAN_PAIR_BURST=2 &
wait
