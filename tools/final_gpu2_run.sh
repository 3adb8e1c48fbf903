#!/bin/bash
# 2-GPU pass (gpurun --gpus 2): multi-rank parity tests (incl. the new Flyer / limiter / device-setup cases) and the
# 2-GPU bench lines with host and device setup
mkdir -p gpurun_out
python -m pytest tests/test_multi_rank.py -m gpu -q > gpurun_out/g2_pytest.log 2>&1; tail -5 gpurun_out/g2_pytest.log
for setup in host device; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    --steps 100 --warmup 10 --setup $setup --no-cpu-baseline > gpurun_out/g2_bench_$setup.log 2>&1
  tail -c 500 gpurun_out/g2_bench_$setup.log; echo
done
