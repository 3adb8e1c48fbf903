#!/bin/bash
# last GPU seconds of round 2: the windowed second density level (fused-step tests, bench line, the 2-rank diagnostic)
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2f3_pytest_fused.log 2>&1; tail -2 $O/r2f3_pytest_fused.log
timeout 60 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/r2f3_bench.log 2>&1; tail -1 $O/r2f3_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac_of_n_gpu_peak'], d['norm_misses_detail']['warmup'], d['clocks'])"
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/diag_norm_miss.py --same-device 1 --real-steps 2 > $O/r2f3_diag.log 2>&1
grep "misses" $O/r2f3_diag.log | cut -c1-120 | tail -6
