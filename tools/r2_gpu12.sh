#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -x -k "two_gpus_match_serial_oracle and p2p and (residual or upwind) or separate_kernels" > $O/r2o_pytest_g2.log 2>&1; tail -4 $O/r2o_pytest_g2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline"
timeout 600 $T > $O/r2o_bench_g2_fused.log 2>&1; tail -1 $O/r2o_bench_g2_fused.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fused', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['parity']['ok'], d['parity']['rhs_relerr'], d['parity']['steps_relerr'], d['clocks'])"
timeout 600 $T --pdl 0 --no-parity > $O/r2o_bench_g2_pdl0.log 2>&1; tail -1 $O/r2o_bench_g2_pdl0.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pdl0', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['clocks'])"
