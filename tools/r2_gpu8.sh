#!/bin/bash
# round 2, 8-GPU pass (gpurun --gpus 8): parity on 4 / 8 ranks, the driver's weak-scaling point (1M points per GPU), BASELINE
# configs[2] (upwind viscosity, 16.8M points, strong scaling point at 8 GPUs) and configs[3] (Sod + residual viscosity, 67M points)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m | head -12 > $O/r2h_topo.txt
line() { python - "$1" "$2" <<'EOF'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'whole %.4f'%d['roofline']['whole_step']['frac_of_n_gpu_peak'], 'setup_s', d['config']['setup_s'], 'points', d['config']['points'], d['parity'], d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
EOF
}
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q -k "many_gpus" > $O/r2h_pytest_g48.log 2>&1; tail -4 $O/r2h_pytest_g48.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --no-cpu-baseline"
timeout 600 $T --steps 100 --warmup 10 > $O/r2h_bench_g8.log 2>&1; line g8_vortex_rv_1M_per_gpu $O/r2h_bench_g8.log
timeout 600 $T --steps 100 --warmup 10 --fused-step 0 --no-parity > $O/r2h_bench_g8_separate.log 2>&1; line g8_separate $O/r2h_bench_g8_separate.log
timeout 900 $T --steps 50 --warmup 5 --source upwind --n-side 1448 --scaling strong > $O/r2h_bench_cfg2_g8_16m.log 2>&1; line cfg2_upwind_16M_g8 $O/r2h_bench_cfg2_g8_16m.log
timeout 1500 $T --steps 30 --warmup 5 --workload sod --n-side 2896 --no-parity > $O/r2h_bench_cfg3_g8_67m.log 2>&1; line cfg3_sod_67M_g8 $O/r2h_bench_cfg3_g8_67m.log
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --no-cpu-baseline"
timeout 600 $T4 --steps 100 --warmup 10 > $O/r2h_bench_g4.log 2>&1; line g4_vortex_rv_1M_per_gpu $O/r2h_bench_g4.log
timeout 900 $T4 --steps 50 --warmup 5 --source upwind --n-side 2048 --scaling strong --no-parity > $O/r2h_bench_cfg2_g4_16m.log 2>&1; line cfg2_upwind_16M_g4 $O/r2h_bench_cfg2_g4_16m.log
T2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --no-cpu-baseline"
timeout 900 $T2 --steps 50 --warmup 5 --source upwind --n-side 2896 --scaling strong --no-parity > $O/r2h_bench_cfg2_g2_16m.log 2>&1; line cfg2_upwind_16M_g2 $O/r2h_bench_cfg2_g2_16m.log
ls -la $O | grep r2h
