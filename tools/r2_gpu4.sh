#!/bin/bash
# round 2, single-GPU pass 3: reworked fused stage kernel (vortex / Sod / upwind), ncu of the stage kernel on Sod, stencil sweep
# on the tile kernels (configs[4]), device setup as the default
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
timeout 900 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2d_pytest_fused.log 2>&1; tail -3 $O/r2d_pytest_fused.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline"
$B > $O/r2d_bench_fused.log 2>&1; line fused $O/r2d_bench_fused.log
$B --workload sod > $O/r2d_bench_fused_sod.log 2>&1; line fused_sod $O/r2d_bench_fused_sod.log
$B --source upwind > $O/r2d_bench_fused_upwind.log 2>&1; line fused_upwind $O/r2d_bench_fused_upwind.log
MFT_LIB_PATH=build/variants/libmft_stage3.so $B > $O/r2d_bench_stage3.log 2>&1; line stage_occ3 $O/r2d_bench_stage3.log
MFT_LIB_PATH=build/variants/libmft_stage3.so $B --workload sod > $O/r2d_bench_stage3_sod.log 2>&1; line stage_occ3_sod $O/r2d_bench_stage3_sod.log
ncu --set full --clock-control none --import-source on -k regex:"k_stage_fused" --launch-skip 6 -c 2 -o $O/r2d_prof_stage_sod -f python bench.py --workload sod --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2d_ncu_sod.log 2>&1
ncu -i $O/r2d_prof_stage_sod.ncu-rep --page raw --csv > $O/r2d_raw_stage_sod.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_stage_fused" --launch-skip 6 -c 2 -o $O/r2d_prof_stage_vortex -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2d_ncu_vortex.log 2>&1
ncu -i $O/r2d_prof_stage_vortex.ncu-rep --page raw --csv > $O/r2d_raw_stage_vortex.csv 2>/dev/null
timeout 1200 python tools/stencil_sweep.py --out $O/r2d_stencil_sweep.json > $O/r2d_stencil_sweep.log 2>&1; tail -8 $O/r2d_stencil_sweep.log
ls -la $O | grep r2d
