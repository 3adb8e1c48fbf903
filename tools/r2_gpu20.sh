#!/bin/bash
# state check after the device layout builder became the default: the whole GPU suite as the driver runs it, smoke(), and the
# bench workload with MFT_TRACE (setup phases) with device-built and host-built layouts
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2w_pytest.log 2>&1; tail -4 $O/r2w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2w_smoke.log 2>&1; tail -1 $O/r2w_smoke.log
B="python bench.py --no-cpu-baseline --steps 100 --warmup 10"
MFT_TRACE=1 $B > $O/r2w_bench_dev.log 2> $O/r2w_bench_dev.err; grep "finalize\|union-tile" $O/r2w_bench_dev.err; tail -1 $O/r2w_bench_dev.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('device layouts', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['config']['setup_s'], d['clocks'])"
MFT_TRACE=1 $B --layout-device 0 > $O/r2w_bench_host.log 2> $O/r2w_bench_host.err; grep "finalize\|union-tile" $O/r2w_bench_host.err; tail -1 $O/r2w_bench_host.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('host layouts', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['config']['setup_s'], d['clocks'])"
