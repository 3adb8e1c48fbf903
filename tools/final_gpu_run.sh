#!/bin/bash
# round-end GPU pass: parity tests, smoke, the default bench line, ncu launch list + full capture of the two pass kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.log 2>gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_tile.csv python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiler --launch-skip 8 -c 2 -o gpurun_out/prof_r1_tile_final -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_r1_tile_final.ncu-rep --page raw --csv > gpurun_out/raw_r1_tile_final.csv 2>/dev/null
ls -la gpurun_out/launches_r1_tile.csv gpurun_out/raw_r1_tile_final.csv
