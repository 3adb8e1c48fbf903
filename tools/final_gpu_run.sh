#!/bin/bash
# GPU pass for the start of the next session (the code of rows f1/f4 has not seen hardware yet):
# parity tests, smoke, the default bench line, setup-pipeline timing, ncu launch list + full capture of the two pass kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; tail -5 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.log 2>gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.log
python bench.py --setup device --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_setup_device.log 2>&1; tail -c 300 gpurun_out/bench_setup_device.log
python tools/setup_bench.py --n-side 1024 > gpurun_out/setup_bench_1m.json 2>gpurun_out/setup_bench.err; cat gpurun_out/setup_bench_1m.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiler --launch-skip 8 -c 2 -o gpurun_out/prof_r2_tile -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_r2_tile.ncu-rep --page raw --csv > gpurun_out/raw_r2_tile.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_setup --launch-skip 2 -c 2 -o gpurun_out/prof_r2_setup -f python tools/setup_bench.py --n-side 512 --reps 1 > gpurun_out/ncu_setup.log 2>&1
ncu -i gpurun_out/prof_r2_setup.ncu-rep --page raw --csv > gpurun_out/raw_r2_setup.csv 2>/dev/null
ls -la gpurun_out/
# named configs at bench size (1 GPU): configs[2]-style upwind viscosity, configs[3]-style Sod + residual viscosity
python bench.py --source upwind --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_vortex_upwind.log 2>&1; tail -c 400 gpurun_out/bench_vortex_upwind.log
python bench.py --workload sod --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_sod_rv.log 2>&1; tail -c 400 gpurun_out/bench_sod_rv.log
bash tools/sanitize.sh
# tuned second copy (layout only, MFT_OPT_TILE bit 4): simulated LDS.128 conflict degree 1.20 -> 1.02
python bench.py --tile 31 --no-cpu-baseline > gpurun_out/bench_tile31.log 2>&1
python -c "import json; d=json.loads(open('gpurun_out/bench_tile31.log').read().strip().splitlines()[-1]); print('tile31', d['value'], d['roofline']['kernel_ms_per_step'])"
ncu --set full --clock-control none --import-source on -k regex:tiler --launch-skip 8 -c 2 -o gpurun_out/prof_r2_tile31 -f python bench.py --tile 31 --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/ncu_full31.log 2>&1
ncu -i gpurun_out/prof_r2_tile31.ncu-rep --page raw --csv > gpurun_out/raw_r2_tile31.csv 2>/dev/null
# rows of a tile ordered by D' row length (pass B: 22.2 -> 20.7 steps per row in the host replay), alone and with the tuned copy
for cfg in "15 1" "31 1"; do
  set -- $cfg
  python bench.py --tile $1 --refine-order $2 --no-cpu-baseline > gpurun_out/bench_tile$1_refine$2.log 2>&1
  python -c "import json; d=json.loads(open('gpurun_out/bench_tile$1_refine$2.log').read().strip().splitlines()[-1]); print('tile $1 refine $2', d['value'], d['roofline']['kernel_ms_per_step'])"
done
# pass B with two rows per thread (same register count as one row per thread: 120 vs 118; 14.3 instead of 22.3 union steps per row in
# the host replay, but 256-row tiles: 60 KB of shared memory per block), default and tuned layout
for t in 15 31; do
  python bench.py --tile $t --tile-rows 21 --no-cpu-baseline > gpurun_out/bench_tile${t}_rowsB2.log 2>&1
  python -c "import json; d=json.loads(open('gpurun_out/bench_tile${t}_rowsB2.log').read().strip().splitlines()[-1]); print('tile $t rows 21', d['value'], d['roofline']['kernel_ms_per_step'])"
done
# occupancy experiment: 6 / 5 resident CTAs per SM for the tile kernels (no spills per ptxas)
for occ in "6 5" "6 4" "5 5"; do
  set -- $occ
  MFT_NVCC_EXTRA="-DMFT_TILE_OCC_A=$1 -DMFT_TILE_OCC_B=$2" python meshfreetrixi.jl_b200/build.py --force > /dev/null 2>&1
  python bench.py --no-cpu-baseline > gpurun_out/bench_occ_$1_$2.log 2>&1
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_occ_$1_$2.log').read().strip().splitlines()[-1]); print('occ $1 $2', d['value'], d['roofline']['kernel_ms_per_step'])"
done
# 256-row tiles
MFT_NVCC_EXTRA="-DMFT_TILE_WARPS=8 -DMFT_TILE_OCC_A=3 -DMFT_TILE_OCC_B=2" python meshfreetrixi.jl_b200/build.py --force > /dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/tile8_pytest.log 2>&1; tail -1 gpurun_out/tile8_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/bench_tile8.log 2>&1
python -c "import json; d=json.loads(open('gpurun_out/bench_tile8.log').read().strip().splitlines()[-1]); print('tile8', d['value'], d['roofline']['kernel_ms_per_step'])"
python meshfreetrixi.jl_b200/build.py --force > /dev/null 2>&1   # back to the default build
