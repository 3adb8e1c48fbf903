#!/bin/bash
# the second density level of the one-pass statistic: the 2-rank diagnostic (expects 0 misses in step 0 now), the fused-step tests
# (incl. the new adjacent-density test) and the multi-rank tests with all ranks on this GPU
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/diag_norm_miss.py --same-device 1 --real-steps 3 > $O/r2z2_diag.log 2>&1
grep "misses\|Error\|error" $O/r2z2_diag.log | cut -c1-260 | tail -8
timeout 600 python -m pytest tests/test_zz_j_fused_step_gpu.py -m gpu -q -x > $O/r2z2_pytest_fused.log 2>&1; tail -3 $O/r2z2_pytest_fused.log
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -x -k "ranks_sharing" > $O/r2z2_pytest_ranks.log 2>&1; tail -3 $O/r2z2_pytest_ranks.log
