#!/bin/bash
mkdir -p gpurun_out
timeout 95 python -m pytest tests/test_multi_rank.py -m gpu -q -x -k "ranks_sharing" > gpurun_out/r2f4_pytest_ranks.log 2>&1; tail -3 gpurun_out/r2f4_pytest_ranks.log
