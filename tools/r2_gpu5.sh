#!/bin/bash
# round 2, single-GPU pass 4: stage kernel with the bulk-async operand ring + redux extremes; multi-rank parity with all ranks on
# one GPU (CUDA IPC between processes, time-sliced); setup timing breakdown
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'EOF'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'whole %.4f'%d['roofline']['whole_step']['frac_of_n_gpu_peak'], 'launches', d['gpu_launches'], 'miss', d.get('norm_misses'), 'setup_s', d['config']['setup_s'], 'e2e %.3g'%d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
EOF
}
nvidia-smi -L
timeout 900 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2e_pytest_fused.log 2>&1; tail -3 $O/r2e_pytest_fused.log
timeout 1500 python -m pytest tests/test_multi_rank.py -m gpu -q -k "sharing_one_gpu" > $O/r2e_pytest_same_device.log 2>&1; tail -12 $O/r2e_pytest_same_device.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline"
MFT_TRACE=1 $B > $O/r2e_bench_fused.log 2>$O/r2e_bench_fused.err; line fused $O/r2e_bench_fused.log; grep -i -E "build|layout|tile|s$" $O/r2e_bench_fused.err | head -20
$B --workload sod > $O/r2e_bench_fused_sod.log 2>&1; line fused_sod $O/r2e_bench_fused_sod.log
$B --source upwind > $O/r2e_bench_fused_upwind.log 2>&1; line fused_upwind $O/r2e_bench_fused_upwind.log
$B --fused-step 0 > $O/r2e_bench_separate.log 2>&1; line separate $O/r2e_bench_separate.log
ncu --set full --clock-control none --import-source on -k regex:"k_stage_fused" --launch-skip 6 -c 2 -o $O/r2e_prof_stage_vortex -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2e_ncu_vortex.log 2>&1
ncu -i $O/r2e_prof_stage_vortex.ncu-rep --page raw --csv > $O/r2e_raw_stage_vortex.csv 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_multi_rank.py > $O/r2e_pytest_all.log 2>&1; tail -5 $O/r2e_pytest_all.log
ls -la $O | grep r2e
