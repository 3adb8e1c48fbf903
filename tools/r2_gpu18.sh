#!/bin/bash
# round 2, second session: (1) fused-step parity tests on the working tree (one GPU: norm combine in block 0 of pass A, no group
# records); (2) one-flag variants against the committed HEAD build, 100 steps each, HEAD measured twice for the noise floor
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k=d['roofline']['kernel_ms_per_step']
    print('%-8s %.4g G  ms/step %.4f  A %.4f B %.4f stage %.4f bc %.4f other %.4f  whole %.4f dom %.4f  miss %s  %s %s' % (sys.argv[1], d['value']/1e9, d['ms_per_step'], k['pass_a'], k['pass_b'], k['stage'], k['bc'], k['other'], d['roofline']['whole_step']['frac_of_n_gpu_peak'], d['roofline']['frac'], d.get('norm_misses'), d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
timeout 600 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2u_pytest_s2.log 2>&1; echo "s1b tests:"; tail -2 $O/r2u_pytest_s2.log
if ! grep -q " passed" $O/r2u_pytest_s2.log || grep -q "failed\|error" $O/r2u_pytest_s2.log; then
  MFT_LIB_PATH=build/variants/libmft_head.so timeout 600 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2u_pytest_s1.log 2>&1; echo "head tests:"; tail -2 $O/r2u_pytest_s1.log
fi
B="python bench.py --no-cpu-baseline --steps 100 --warmup 10"
for v in head s1b s1b_b8o4 s1b_b8o5 s1b_b6o5 head s1b s1b_a8o5 s1b_a6o6 s1b_a2o7; do
  MFT_LIB_PATH=build/variants/libmft_$v.so timeout 300 $B > $O/r2u_bench_$v.log 2>&1; line $v $O/r2u_bench_$v.log
done
