"""Compare the SASS instruction streams of two builds of libmft_b200.so kernel by kernel (addresses and encodings stripped).

    python tools/sass_compare.py old.so [new.so]      # new.so defaults to the in-tree library

Used to show that host-side refactors and additions leave the measured kernels untouched: build the last hardware-measured
commit into a scratch directory (`git archive <commit> meshfreetrixi.jl_b200/csrc include | tar -x -C /tmp/old` and the nvcc
line of meshfreetrixi.jl_b200/build.py), then compare.  Needs cuobjdump and c++filt; no GPU."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fns, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fns[cur] = []
            continue
        mm = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);\s*/\*", line)
        if cur is not None and mm:
            fns[cur].append(mm.group(1).strip())
    return fns


def main():
    old = kernels(sys.argv[1])
    new = kernels(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "meshfreetrixi.jl_b200", "libmft_b200.so"))
    common = [k for k in old if k in new]
    changed = [k for k in common if old[k] != new[k]]
    print(f"{len(old)} kernels in the old build, {len(new)} in the new one, {len(common)} in both, "
          f"{len(common) - len(changed)} with identical instruction streams")
    names = lambda ks: subprocess.run(["c++filt"] + list(ks), capture_output=True, text=True).stdout.splitlines() if ks else []
    for label, ks in (("changed", changed), ("removed", [k for k in old if k not in new]), ("added", [k for k in new if k not in old])):
        for n in names(ks):
            print(f"  {label}: {n[:140]}")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
