#!/bin/bash
# Diagnostic variants of the library (NOT shipped; most compute garbage by construction): which resource binds the
# stream-K layer-1 products?  Load one with WGAN_B200_LIB=<file name> (see _lib.py).
#   nob: stream A only   nomma: 1 of 4 MMAs per K block   swc: no tensor work, slots freed by plain arrives
#   noepi: accumulators dropped   burstN: producer issues K blocks in groups of N
cd "$(dirname "$0")/.." || exit 1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
S=arshamg_scrnaseq_wgan_b200/csrc/engine.cu
O=arshamg_scrnaseq_wgan_b200
build() { name=$1; shift; nvcc $F "$@" -o $O/libwgan_diag_$name.so $S 2>&1 | grep -E "error" ; }
build stamps -DWGAN_DIAG_STAMPS &
wait
