#!/bin/bash
# round 2, 8-GPU pass 2: the driver's weak-scaling points (1M points per GPU) at 8 and 4 GPUs with the final multi-GPU code
mkdir -p gpurun_out
O=gpurun_out
line() { python - "$1" "$2" <<'EOF'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms_per_step'], 'whole %.4f'%d['roofline']['whole_step']['frac_of_n_gpu_peak'], 'setup_s', d['config']['setup_s'], 'points', d['config']['points'], d['parity'], d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
EOF
}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --no-cpu-baseline"
timeout 600 $T --steps 100 --warmup 10 > $O/r2p_bench_g8.log 2>&1; line g8_vortex_rv_1M_per_gpu $O/r2p_bench_g8.log
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --no-cpu-baseline"
timeout 600 $T4 --steps 100 --warmup 10 > $O/r2p_bench_g4.log 2>&1; line g4_vortex_rv_1M_per_gpu $O/r2p_bench_g4.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/r2p_bench_g1.log 2>&1; line g1_same_box $O/r2p_bench_g1.log
