#!/bin/bash
# round 2, final single-GPU pass of the second session: the whole GPU test suite as the driver runs it, smoke(), the default bench
# line (with the CPU baseline), the fma-mode line for the record, ncu launch list + full capture of the three kernels of a fused stage
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2f2_pytest.log 2>&1; tail -3 $O/r2f2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f2_smoke.log 2>&1; tail -1 $O/r2f2_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2f2_bench_driver_cmd.log 2>$O/r2f2_bench_driver_cmd.err; tail -1 $O/r2f2_bench_driver_cmd.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('driver cmd', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac_of_n_gpu_peak'], 'e2e %.4g'%d['e2e']['value'], d['norm_misses_detail'], d['clocks'], d['config']['setup_s'])"
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/r2f2_bench_200.log 2>&1; tail -1 $O/r2f2_bench_200.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('200 steps', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac_of_n_gpu_peak'], d['clocks'])"
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --fma > $O/r2f2_bench_fma.log 2>&1; tail -1 $O/r2f2_bench_fma.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fma mode', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac_of_n_gpu_peak'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2f2_launches.csv python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2f2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tiler|k_stage_fused" --launch-skip 12 -c 3 -o $O/r2f2_prof -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2f2_ncu_full.log 2>&1
ncu -i $O/r2f2_prof.ncu-rep --page raw --csv > $O/r2f2_raw.csv 2>/dev/null
ls -la $O | grep r2f2 | awk '{print $5, $9}'
