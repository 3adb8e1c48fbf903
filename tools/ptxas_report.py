#!/usr/bin/env python
"""Registers / spills of the kernels of a fused stage from a `-Xptxas -v` log (build/variants/<name>.ptxas.log)."""
import re, subprocess, sys

KEEP = ('k_pass_a_tiler<1, true, true, 2, false>', 'k_pass_a_tiler<1, true, true, 1, false>', 'k_pass_b_tiler<1, true, false>',
        'k_stage_fused<1, false>', 'k_stage_fused<1, true>')
for path in sys.argv[1:]:
    txt = open(path).read()
    ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                      r"ptxas info\s*:\s*Used (\d+) registers", txt)
    want = [e for e in ents if any(k in e[0] for k in ('k_pass_a_tiler', 'k_pass_b_tiler', 'k_stage_fused'))]
    dem = subprocess.run(['cu++filt'] + [w[0] for w in want], capture_output=True, text=True).stdout.strip().splitlines()
    for d, (m, stack, ss, sl, regs) in zip(dem, want):
        d = d.replace('void mft::', '').replace('(int)', '').replace('(bool)1', 'true').replace('(bool)0', 'false')
        d = re.sub(r'>\(.*', '>', d)
        if d in KEEP:
            print(f'{path.split("/")[-1]:24s} {d:42s} regs {regs:>3s} stack {stack} spill {ss}/{sl}')
