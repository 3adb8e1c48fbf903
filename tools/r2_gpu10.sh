#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_pass_a_tiler" --launch-skip 6 -c 2 -o $O/r2l_prof_passa -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2l_ncu.log 2>&1
ncu -i $O/r2l_prof_passa.ncu-rep --page raw --csv > $O/r2l_raw_passa.csv 2>/dev/null
ls -la $O | grep r2l
