#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'EOF'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'whole %.4f'%d['roofline']['whole_step']['frac_of_n_gpu_peak'], 'dom %.4f'%d['roofline']['frac'], 'miss', d.get('norm_misses'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
EOF
}
timeout 600 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2r_pytest.log 2>&1; tail -2 $O/r2r_pytest.log
B="python bench.py --no-cpu-baseline --steps 100 --warmup 10"
$B > $O/r2r_bench_a.log 2>&1; line default_a $O/r2r_bench_a.log
MFT_LIB_PATH=build/variants/libmft_occ75.so $B > $O/r2r_bench_occ75.log 2>&1; line occ75 $O/r2r_bench_occ75.log
$B > $O/r2r_bench_b.log 2>&1; line default_b $O/r2r_bench_b.log
MFT_LIB_PATH=build/variants/libmft_occ75.so $B > $O/r2r_bench_occ75b.log 2>&1; line occ75_b $O/r2r_bench_occ75b.log
