#!/usr/bin/env python
"""Refresh profiles/traffic.json: DRAM bytes per launch of the step kernel (ncu dram__bytes_read.sum + dram__bytes_write.sum) for the
bench workloads, stamped with the hash of the kernel sources it was taken from.  bench.py reports `roofline.traffic` from this file
ONLY while that hash equals the current sources, so a changed kernel can never carry a stale figure.

    gpurun -- 'python scripts/refresh_traffic.py'        # one GPU; then commit profiles/traffic.json

Launches 11..12 of the step kernel of each bench command are captured (after the reset launch and the warm-up), two per workload.
"""
import csv
import io
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

WORK = [  # (key fields, bench flags)
    (("c2", 4096, 4), []),
    (("c2", 8192, 4), ["--envs-per-gpu", "8192"]),
    (("c3", 16384, 4), ["--workload", "c3"]),
    (("c2", 4096, 20), ["--substeps", "20"]),
]


def capture(flags):
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "--kernel-name-base", "demangled", "-k", r"regex:zb_kernel<\(int\)0", "-s", "10", "-c", "2", "--csv",
           sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "10", "--no-cpu-baseline", "--no-extras", *flags]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout
    rows = [r for r in csv.reader(io.StringIO(out[out.find('"ID"'):])) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    name, val, unit, kern = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    per = {}
    for r in data:
        v = float(r[val].replace(",", ""))
        u = r[unit].lower()
        if r[name].startswith("dram__bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        elif u.startswith("us"):
            v *= 1e-3
        elif u.startswith("ns"):
            v *= 1e-6
        per.setdefault(r[0], {"kernel": r[kern][:40]})[r[name]] = v
    ids = sorted(per)
    tot = [per[i]["dram__bytes_read.sum"] + per[i]["dram__bytes_write.sum"] for i in ids]
    return {"dram_bytes_per_launch": sum(tot) / len(tot), "read": sum(per[i]["dram__bytes_read.sum"] for i in ids) / len(ids),
            "write": sum(per[i]["dram__bytes_write.sum"] for i in ids) / len(ids), "launches_captured": len(ids), "kernel": per[ids[0]]["kernel"],
            "ms_under_ncu": sum(per[i].get("gpu__time_duration.sum", 0.0) for i in ids) / len(ids)}


def main():
    caps = {}
    for (w, n, s), flags in WORK:
        for obs in (True, False):
            if not obs and (w, n, s) != ("c2", 4096, 4):
                continue
            key = f"{w}_n{n}_s{s}_{'obs' if obs else 'noobs'}_cg"
            try:
                caps[key] = capture(flags + ([] if obs else ["--no-obs"]))
                rnd = lambda x: bench.algorithmic_bytes_per_ctrl_step(w != "c2", w != "c2", obs) * n  # noqa: E731
                caps[key]["algorithmic_bytes_per_launch"] = rnd(0)
                print(key, caps[key], flush=True)
            except Exception as ex:  # noqa: BLE001
                print(key, "capture failed:", ex, flush=True)
    out = {"kernel_source_sha": bench.kernel_source_sha(), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, launches 11-12 of zb_kernel<0,...> of `bench.py --steps 4 --warmup 10`",
           "captures": caps}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for p in (os.path.join(ROOT, "gpurun_out", "traffic.json"), os.path.join(ROOT, "profiles", "traffic.json")):
        json.dump(out, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
