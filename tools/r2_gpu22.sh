#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/diag_norm_miss.py --same-device 1 > gpurun_out/r2y_diag.log 2>&1
grep -v "Warning\|warn\|^$\|OMP_NUM\|\*\*\*\*" gpurun_out/r2y_diag.log | tail -25
