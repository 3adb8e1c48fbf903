#!/bin/bash
# compute-sanitizer pass over the small fixture (the reference has no race detection / sanitizers).
# memcheck + racecheck on smoke() (Euler + residual viscosity rhs! and two graph-replayed SSPRK33 steps, 2154 points) and
# on the setup / limiter / IGR GPU tests.  Run on a GPU box:  bash tools/sanitize.sh   -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck; do
  $SAN --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "$tool smoke: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_${tool}_smoke.log
done
$SAN --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_g_setup_gpu.py tests/test_zz_f_limiter.py tests/test_zz_i_igr_gpu.py tests/test_zz_b_aux_gpu.py \
  -m gpu -q -k "not million" > gpurun_out/sanitize_memcheck_next_rows.log 2>&1
echo "memcheck next rows: exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_next_rows.log
