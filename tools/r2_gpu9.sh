#!/bin/bash
# round 2, single-GPU pass 6: norm finalize moved into block 0 of pass A (one GPU), programmatic dependent launch between the
# kernels of a fused stage, single-fence flag stores; fused tests + ranks-on-one-GPU tests + bench with / without PDL
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
timeout 900 python -m pytest tests/test_zz_j_fused_step_gpu.py tests/test_multi_rank.py tests/test_zz_e_time_dependent_bc_gpu.py -m gpu -q -x -k "fused or sharing_one_gpu or time_dependent" > $O/r2i_pytest_fused.log 2>&1; tail -5 $O/r2i_pytest_fused.log
B="python bench.py --no-cpu-baseline --steps 100 --warmup 10"
$B > $O/r2i_bench_pdl1.log 2>&1; line pdl1 $O/r2i_bench_pdl1.log
$B --pdl 0 > $O/r2i_bench_pdl0.log 2>&1; line pdl0 $O/r2i_bench_pdl0.log
$B --workload sod > $O/r2i_bench_sod.log 2>&1; line sod $O/r2i_bench_sod.log
$B --source upwind > $O/r2i_bench_upwind.log 2>&1; line upwind $O/r2i_bench_upwind.log
$B --graph 0 > $O/r2i_bench_eager.log 2>&1; line eager $O/r2i_bench_eager.log
ls -la $O | grep r2i
