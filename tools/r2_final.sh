#!/bin/bash
# round 2, final single-GPU pass: the whole GPU test suite as the driver runs it, smoke(), the default bench line (with the CPU
# baseline), ncu launch list + full capture of the three kernels of a fused stage (-> profiles/r2_final_*, traffic.json)
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2z_pytest.log 2>&1; tail -4 $O/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1; tail -1 $O/r2z_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2z_bench_driver_cmd.log 2>$O/r2z_bench_driver_cmd.err; tail -c 1200 $O/r2z_bench_driver_cmd.log; echo
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/r2z_bench_200.log 2>&1; tail -1 $O/r2z_bench_200.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('200 steps', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['whole_step'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2z_launches.csv python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2z_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tiler|k_stage_fused" --launch-skip 12 -c 3 -o $O/r2z_prof -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2z_ncu_full.log 2>&1
ncu -i $O/r2z_prof.ncu-rep --page raw --csv > $O/r2z_raw.csv 2>/dev/null
ls -la $O | grep r2z
