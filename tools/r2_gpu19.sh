#!/bin/bash
# device-built union-tile layouts: byte identity with the host builder + build times (MFT_TRACE), then the bench setup time both ways
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_zz_k_device_layout_gpu.py -m gpu -q -x -s > $O/r2v_pytest_layout.log 2>&1; tail -25 $O/r2v_pytest_layout.log
