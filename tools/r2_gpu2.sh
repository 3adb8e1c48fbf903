#!/bin/bash
# round 2, second single-GPU pass: first hardware run of the fused step (k_stage_fused + norm statistic), full parity suite,
# fused vs separate kernels in the bench, ncu launch list + full capture of the fused stage kernel and the tile kernels
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'E'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'whole %.4f'%d['roofline']['whole_step']['frac_of_n_gpu_peak'], 'launches', d['gpu_launches'], 'miss', d.get('norm_misses'), 'e2e %.3g'%d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
E
}
timeout 900 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2b_pytest_fused.log 2>&1; tail -15 $O/r2b_pytest_fused.log
timeout 900 python -m pytest tests -m gpu -q > $O/r2b_pytest.log 2>&1; tail -8 $O/r2b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2b_smoke.log 2>&1; tail -1 $O/r2b_smoke.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline"
$B > $O/r2b_bench_fused.log 2>&1; line fused $O/r2b_bench_fused.log
$B --fused-step 0 > $O/r2b_bench_separate.log 2>&1; line separate $O/r2b_bench_separate.log
$B --source upwind > $O/r2b_bench_fused_upwind.log 2>&1; line fused_upwind $O/r2b_bench_fused_upwind.log
$B --workload sod > $O/r2b_bench_fused_sod.log 2>&1; line fused_sod $O/r2b_bench_fused_sod.log
for v in occ75 occ64; do
  MFT_LIB_PATH=build/variants/libmft_$v.so $B > $O/r2b_bench_$v.log 2>&1; line $v $O/r2b_bench_$v.log
done
python bench.py > $O/r2b_bench_default_full.log 2>$O/r2b_bench_default_full.err; line default_full $O/r2b_bench_default_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2b_launches.csv python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2b_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tiler|k_stage_fused" --launch-skip 12 -c 3 -o $O/r2b_prof -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2b_ncu_full.log 2>&1
ncu -i $O/r2b_prof.ncu-rep --page raw --csv > $O/r2b_raw.csv 2>/dev/null
ls -la $O | grep r2b
