#!/bin/bash
# round 2, 2-GPU pass (gpurun --gpus 2): multi-rank parity (fused step with band-tile waits + fused puts, and the separate
# kernels), then the 2-GPU bench line with the parity block, fused vs separate
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_multi_rank.py -m gpu -q -x > $O/r2c_pytest_g2.log 2>&1; tail -15 $O/r2c_pytest_g2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline"
timeout 600 $T > $O/r2c_bench_g2_fused.log 2>&1; tail -c 1500 $O/r2c_bench_g2_fused.log; echo
timeout 600 $T --fused-step 0 > $O/r2c_bench_g2_separate.log 2>&1; tail -c 1500 $O/r2c_bench_g2_separate.log; echo
