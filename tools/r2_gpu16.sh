#!/bin/bash
mkdir -p gpurun_out
MFT_LIB_PATH=build/variants/libmft_stagedbg.so python bench.py --no-cpu-baseline --steps 4 --warmup 3 --graph 0 > gpurun_out/r2s_stagedbg.log 2>&1
grep "stage dbg" gpurun_out/r2s_stagedbg.log | tail -12
