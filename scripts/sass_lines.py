"""Static SASS instruction count per CUDA source line of one kernel, from an object file built with -lineinfo (no GPU needed).
The L1.5 instruction cache of an SM holds 32 KB = 2048 instructions: this shows where a kernel's code size comes from.

    python scripts/sass_lines.py rayz_b200/lib/obj/rz_path.o rz_second_kernelILb0 [top]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
cnt, byfile, cur, on = collections.Counter(), collections.Counter(), None, False
for line in out.splitlines():
    if line.startswith(".text."):
        on = pat in line
        continue
    if line.startswith(".section") or (line.startswith(".") and not line.startswith(".L")):
        if not line.startswith(".text."):
            pass
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        cnt[cur] += 1
        byfile[cur[0]] += 1
tot = sum(cnt.values())
print(f"{tot} instructions = {tot * 16 / 1024:.1f} KB;", dict(byfile))
for k, v in cnt.most_common(top):
    print(f"{k[0]}:{k[1]:<6d} {v}")
