#!/bin/bash
# build/variants/libmft_<name>.so from the working tree with extra nvcc flags (experiment knobs); prints registers / spills of the
# three kernels of a fused stage.   usage: tools/build_variant.sh <name> [extra nvcc flags...]
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -shared -Xcompiler -fPIC -Xcompiler -O2 \
     -Xcompiler -pthread -cudart static -Xptxas -v "$@" -o build/variants/libmft_$name.so meshfreetrixi.jl_b200/csrc/mft_b200.cu -ldl \
     > build/variants/$name.ptxas.log 2>&1 || { tail -20 build/variants/$name.ptxas.log; exit 1; }
python tools/ptxas_report.py build/variants/$name.ptxas.log
