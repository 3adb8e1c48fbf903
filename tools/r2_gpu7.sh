#!/bin/bash
# round 2, single-GPU pass 5: validation of the per-warp ring / warp-cooperative finalize build, then the 1-GPU denominators of
# the named configs: configs[2] (upwind viscosity, 16.8M points) and configs[3] (Sod + RV, 8.4M points = one GPU's share of 67M)
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'EOF'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'whole %.4f'%d['roofline']['whole_step']['frac_of_n_gpu_peak'], 'launches', d['gpu_launches'], 'miss', d.get('norm_misses'), 'setup_s', d['config']['setup_s'], 'points', d['config']['points'], 'e2e %.3g'%d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
EOF
}
timeout 900 python -m pytest tests/test_zz_j_fused_step_gpu.py tests/test_multi_rank.py -m gpu -q -x -k "fused or sharing_one_gpu" > $O/r2g_pytest_fused.log 2>&1; tail -3 $O/r2g_pytest_fused.log
B="python bench.py --no-cpu-baseline"
$B --steps 100 --warmup 10 > $O/r2g_bench_fused.log 2>&1; line fused $O/r2g_bench_fused.log
$B --steps 100 --warmup 10 --workload sod > $O/r2g_bench_fused_sod.log 2>&1; line fused_sod $O/r2g_bench_fused_sod.log
$B --steps 100 --warmup 10 --source upwind > $O/r2g_bench_fused_upwind.log 2>&1; line fused_upwind $O/r2g_bench_fused_upwind.log
$B --steps 30 --warmup 5 --source upwind --n-side 4096 --scaling strong > $O/r2g_bench_cfg2_g1_16m.log 2>&1; line cfg2_upwind_16M_g1 $O/r2g_bench_cfg2_g1_16m.log
$B --steps 30 --warmup 5 --workload sod --n-side 2896 > $O/r2g_bench_cfg3_g1_8m.log 2>&1; line cfg3_sod_8M_g1 $O/r2g_bench_cfg3_g1_8m.log
ncu --set full --clock-control none --import-source on -k regex:"k_stage_fused" --launch-skip 6 -c 2 -o $O/r2g_prof_stage_vortex -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2g_ncu_vortex.log 2>&1
ncu -i $O/r2g_prof_stage_vortex.ncu-rep --page raw --csv > $O/r2g_raw_stage_vortex.csv 2>/dev/null
ls -la $O | grep r2g
