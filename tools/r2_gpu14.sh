#!/bin/bash
# round 2: L2 prefetch distance sweep of the tile kernels (the resident block count grew from 4 to 6 / 5 per SM this round), and the
# reference arm (CPU, full cloud) as the driver will run it
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'EOF'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'whole %.4f'%d['roofline']['whole_step']['frac_of_n_gpu_peak'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
EOF
}
B="python bench.py --no-cpu-baseline --steps 100 --warmup 10"
for pf in 592 1184 2368 3552 4736 7104; do
  $B --pf-dist $pf > $O/r2q_bench_pf$pf.log 2>&1; line pf$pf $O/r2q_bench_pf$pf.log
done
( time python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r2q_reference.log 2>&1; tail -5 $O/r2q_reference.log | cut -c1-600
