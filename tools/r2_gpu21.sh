#!/bin/bash
# two GPUs: the multi-GPU parity tests and the bench line (with its parity block) on the current defaults (device-built layouts)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -x > $O/r2x_pytest_g2.log 2>&1; tail -3 $O/r2x_pytest_g2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > $O/r2x_bench_g2.log 2> $O/r2x_bench_g2.err
tail -1 $O/r2x_bench_g2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2 GPUs', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d.get('parity'), d['config']['setup_s'], d['clocks'], 'misses', d.get('norm_misses'))" || tail -5 $O/r2x_bench_g2.err
