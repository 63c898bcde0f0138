#!/usr/bin/env python
"""Warp-stall samples and executed instructions of the step kernel by source function and by source line, from the
`--page source` view of an `ncu --set full --import-source on` capture (the kernels are compiled with -lineinfo).
usage: python profiles/ncu_hotspots.py gpurun_out/prof.ncu-rep [n_lines]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, lines, calls = None, None, [], []
srcs = collections.defaultdict(dict)
last = None
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        last = (cur, int(r[0]), r[1].strip())
        lines.append((cur, int(r[0]), r[1].strip(), d))
        srcs[cur][int(r[0])] = r[1]
    elif hdr and len(r) == len(hdr) and r[2].startswith("0x") and "CALL" in r[3]:
        calls.append((last, r[3].strip(), r[7]))
num = lambda d, k: int(d[k]) if d.get(k, "").isdigit() else 0
tot = sum(num(d, "# Samples") for *_, d in lines) or 1
ti = sum(num(d, "Instructions Executed") for *_, d in lines) or 1
# function of a line: nearest preceding function header in the same file (read from the working tree: the report's
# source view only lists lines that own instructions)
import os
fstart = collections.defaultdict(list)
csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ksim-gym-zbot_b200", "csrc")
for f in srcs:
    path = os.path.join(csrc, f)
    if not os.path.exists(path):
        continue
    for ln, text in enumerate(open(path).read().split("\n"), 1):
        g = re.match(r"\s*(?:static )?(?:ZB_DI|__device__ __noinline__|__global__ void)[^(;=]*?(\w+)\(", text)
        if g:
            fstart[f].append((ln, g.group(1)))
def fn(f, ln):
    name = "-"
    for s, n in fstart.get(f, []):
        if s <= ln:
            name = n
    return name
byf, byi = collections.Counter(), collections.Counter()
for f, ln, s, d in lines:
    k = (f, fn(f, ln))
    byf[k] += num(d, "# Samples"); byi[k] += num(d, "Instructions Executed")
print(f"total samples {tot}, warp instructions executed {ti}")
print("\n== by function (file, function, % samples, % instructions) ==")
for k, v in byf.most_common(30):
    print(f"{k[0][:20]:20s} {k[1]:18s} {100 * v / tot:5.1f}%  {100 * byi[k] / ti:5.1f}%")
print("\n== by line ==")
for f, ln, s, d in sorted(lines, key=lambda x: -num(x[3], "# Samples"))[:nlines]:
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not" not in k and v.isdigit() and int(v) > 0}
    top = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{f[:14]:14s}{ln:5d} {100 * num(d, '# Samples') / tot:4.1f}% i{100 * num(d, 'Instructions Executed') / ti:4.1f}%  {s[:84]:84s} {top}")
print("\n== out-of-line calls that executed ==")
for (f, ln, s), ins, ex in calls:
    if ex.isdigit() and int(ex) > 0:
        print(f"{f}:{ln}  {s[:70]}  | {ins}  x{ex}")
