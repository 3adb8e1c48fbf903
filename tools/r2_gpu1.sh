#!/bin/bash
# round 2, first single-GPU pass: parity tests, the default bench line, the queued one-flag kernel experiments (prebuilt variant
# libraries in build/variants/, selected with MFT_LIB_PATH), configs[2]/[3] workloads at 1M points, setup timing, sanitizer, ncu
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'E'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'e2e %.3g'%d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
E
}
python -m pytest tests -m gpu -q -x > $O/r2a_pytest.log 2>&1; tail -3 $O/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1; tail -1 $O/r2a_smoke.log
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline"
$B > $O/r2a_bench_default.log 2>&1; line default $O/r2a_bench_default.log
$B --tile 31 > $O/r2a_bench_tile31.log 2>&1; line tile31 $O/r2a_bench_tile31.log
$B --tile 15 --refine-order 1 > $O/r2a_bench_refine.log 2>&1; line tile15_refine $O/r2a_bench_refine.log
$B --tile 31 --refine-order 1 > $O/r2a_bench_tile31_refine.log 2>&1; line tile31_refine $O/r2a_bench_tile31_refine.log
$B --tile 15 --tile-rows 21 > $O/r2a_bench_rowsB2.log 2>&1; line tile15_rowsB2 $O/r2a_bench_rowsB2.log
$B --tile 31 --tile-rows 21 > $O/r2a_bench_tile31_rowsB2.log 2>&1; line tile31_rowsB2 $O/r2a_bench_tile31_rowsB2.log
for v in occ65 occ55; do
  MFT_LIB_PATH=build/variants/libmft_$v.so $B > $O/r2a_bench_$v.log 2>&1; line $v $O/r2a_bench_$v.log
  MFT_LIB_PATH=build/variants/libmft_$v.so $B --tile 31 > $O/r2a_bench_${v}_tile31.log 2>&1; line ${v}_tile31 $O/r2a_bench_${v}_tile31.log
done
MFT_LIB_PATH=build/variants/libmft_tile8.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r2a_tile8_pytest.log 2>&1; tail -1 $O/r2a_tile8_pytest.log
MFT_LIB_PATH=build/variants/libmft_tile8.so $B > $O/r2a_bench_tile8.log 2>&1; line tile8 $O/r2a_bench_tile8.log
MFT_LIB_PATH=build/variants/libmft_tile8.so $B --tile 31 > $O/r2a_bench_tile8_tile31.log 2>&1; line tile8_tile31 $O/r2a_bench_tile8_tile31.log
# named configs at 1M points
$B --source upwind > $O/r2a_bench_vortex_upwind.log 2>&1; line upwind $O/r2a_bench_vortex_upwind.log
$B --workload sod > $O/r2a_bench_sod_rv.log 2>&1; line sod $O/r2a_bench_sod_rv.log
$B --setup device > $O/r2a_bench_setup_device.log 2>&1; line setup_device $O/r2a_bench_setup_device.log
python tools/setup_bench.py --n-side 1024 > $O/r2a_setup_bench_1m.json 2>$O/r2a_setup_bench.err; cat $O/r2a_setup_bench_1m.json
# ncu: launch list + full capture of the two pass kernels (default and tuned layout)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches.csv python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2a_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tiler --launch-skip 8 -c 2 -o $O/r2a_prof_tile -f python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2a_ncu_full.log 2>&1
ncu -i $O/r2a_prof_tile.ncu-rep --page raw --csv > $O/r2a_raw_tile.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:tiler --launch-skip 8 -c 2 -o $O/r2a_prof_tile31 -f python bench.py --tile 31 --steps 2 --warmup 3 --graph 0 --no-cpu-baseline > $O/r2a_ncu_full31.log 2>&1
ncu -i $O/r2a_prof_tile31.ncu-rep --page raw --csv > $O/r2a_raw_tile31.csv 2>/dev/null
# sanitizer (small fixture)
timeout 600 bash tools/sanitize.sh
ls -la $O | head -60
