"""Static SASS instruction count and executed warp instructions per CUDA source line from an ncu report (code-size accounting:
the L1.5 instruction cache holds 32 KB = 2048 instructions, so which source lines the kernel's instructions come from matters).

    python scripts/sass_by_line.py gpurun_out/x.ncu-rep [top]
"""
import csv
import collections
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
static = collections.Counter()
dyn = collections.Counter()
samples = collections.Counter()
text = {}
cur_file, cur = "?", None
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_addr, i_exec, i_samp = r.index("Address"), r.index("Instructions Executed"), r.index("# Samples")
        continue
    if hdr is None or len(r) <= i_exec:
        continue
    if r[0]:
        cur = (cur_file, int(r[0]))
        text[cur] = r[1].strip()[:110]
    elif r[i_addr].startswith("0x") and cur:
        static[cur] += 1
        dyn[cur] += int(r[i_exec] or 0)
        samples[cur] += int(r[i_samp] or 0)
tot_s, tot_d, tot_p = sum(static.values()), sum(dyn.values()), sum(samples.values())
print(f"static SASS instructions {tot_s} ({tot_s * 16 / 1024:.1f} KB), executed warp instructions {tot_d}, samples {tot_p}")
byfile = collections.Counter()
for k, v in static.items():
    byfile[k[0]] += v
print("by file:", dict(byfile))
print(f"{'file:line':28s} {'static':>6s} {'exec %':>7s} {'samp %':>7s}  source")
order = samples.most_common(top) if "--by-samples" in sys.argv else dyn.most_common(top) if "--by-exec" in sys.argv else static.most_common(top)
for k, _ in order:
    print(f"{k[0] + ':' + str(k[1]):28s} {static[k]:6d} {100 * dyn[k] / tot_d:7.2f} {100 * samples[k] / tot_p:7.2f}  {text.get(k, '')}")
# executed instructions / samples by file, and by 50-line bands of each file (which part of the kernel the time goes to)
bands = collections.Counter()
bands_p = collections.Counter()
for k in static:
    b = (k[0], k[1] // 50 * 50)
    bands[b] += dyn[k]
    bands_p[b] += samples[k]
print("\nby 50-line band (exec % / samples %), bands above 0.5 % of either:")
for b in sorted(bands):
    e, p = 100 * bands[b] / tot_d, 100 * bands_p[b] / tot_p
    if e >= 0.5 or p >= 0.5:
        print(f"  {b[0]}:{b[1]}-{b[1] + 49}  {e:6.2f} {p:6.2f}")
