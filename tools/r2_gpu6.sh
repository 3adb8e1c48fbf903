#!/bin/bash
# round 2, 2-GPU pass 2 (gpurun --gpus 2): multi-rank parity on two devices (fused + separate kernels, both data planes),
# 2-GPU bench lines with the parity block and the box topology
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L; nvidia-smi topo -m | head -8
timeout 1500 python -m pytest tests/test_multi_rank.py -m gpu -q > $O/r2f_pytest_g2.log 2>&1; tail -8 $O/r2f_pytest_g2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline"
timeout 600 $T > $O/r2f_bench_g2_fused.log 2>&1; tail -1 $O/r2f_bench_g2_fused.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fused', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['parity'], d['config']['gpu_topology'], d['config']['setup_s'])"
timeout 600 $T --fused-step 0 > $O/r2f_bench_g2_separate.log 2>&1; tail -1 $O/r2f_bench_g2_separate.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('separate', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
timeout 600 $T --source upwind > $O/r2f_bench_g2_upwind.log 2>&1; tail -1 $O/r2f_bench_g2_upwind.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('upwind', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['parity'])"
