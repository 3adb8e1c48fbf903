"""bench.py's handling of norm misses, driven on the CPU tier through the dry-run stand-in (tests/dryrun/mocklib.py: the oracle behind
the C-ABI entry points; it says nothing about the CUDA code).  What is checked is the bench's own logic: a region that had a miss
is never the one that is timed -- misses during warm-up are reported and the run goes on, misses inside the timed region send the
context to the two-pass kernels (MFT_OPT_FUSED_STEP = 0) and the region is run again -- and the JSON line stays well formed."""
import json
import os
import subprocess
import sys

import cases

CODE = ("import sys, runpy; sys.path.insert(0, %r); import mocklib; mocklib.install(); "
        "sys.argv = ['bench.py', '--n-side', '24', '--steps', '2', '--warmup', '3', '--no-cpu-baseline', '--setup', 'host']; "
        "runpy.run_path(%r, run_name='__main__')") % (os.path.join(cases.ROOT, "tests", "dryrun"), os.path.join(cases.ROOT, "bench.py"))


def _bench(misses=None):
    env = dict(os.environ)
    env.pop("MFT_MOCK_NORM_MISSES", None)
    if misses:
        env["MFT_MOCK_NORM_MISSES"] = misses
    res = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=600, cwd=cases.ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1]), res.stderr


def test_bench_line_without_misses():
    d, _ = _bench()
    assert d["norm_misses"] == 0 and d["norm_misses_detail"] == dict(d["norm_misses_detail"], warmup=0, timed=0, fallback=None)
    assert d["config"]["layout"]["fused_step"] == 1
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key


def test_misses_during_warmup_are_reported_and_the_fused_step_is_kept():
    d, err = _bench("2,7")          # the 2nd of 3 warm-up steps misses 7 rows
    assert d["norm_misses_detail"]["warmup"] == 7 and d["norm_misses_detail"]["fallback"] is None
    assert d["config"]["layout"]["fused_step"] == 1 and d["norm_misses"] == 7
    assert "re-running" not in err


def test_misses_inside_the_timed_region_fall_back_to_the_two_pass_kernels():
    d, err = _bench("4,3")          # the 1st timed step (4th since the upload) misses 3 rows
    assert "re-running with the two-pass kernels" in err
    assert d["norm_misses_detail"]["fallback"] and d["norm_misses_detail"]["timed"] == 0
    assert d["config"]["layout"]["fused_step"] == 0
    assert d["norm_misses"] == 3   # the counter keeps what the first region saw; the timed one added nothing
