#!/usr/bin/env python
"""Static SASS statistics per kernel of a built library / cubin: instruction count, spills (LDL / STL), shuffles, selects, WARPSYNC
(a non-zero WARPSYNC count in a production instantiation means ptxas could not prove the control flow around the shuffles warp-uniform).
usage: python scripts/sass_stats.py [lib.so ...]"""
import collections, re, subprocess, sys, os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:] or [os.path.join(ROOT, "ksim-gym-zbot_b200", "csrc", "libzbphys.so")]
KEYS = ["LDL", "STL", "SHFL", "FFMA", "FMUL", "FADD", "FSEL", "SEL", "LOP3", "ISETP", "FSETP", "MUFU", "LDS", "STS", "LDG", "BAR", "BSSY", "WARPSYNC", "VOTE", "CALL"]
for lib in libs:
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    print(lib)
    for p in re.split(r"\n\s*Function : ", txt)[1:]:
        name = p.split("\n")[0]
        g = re.match(r"_Z9zb_kernelILi(\d)ELb(\d)ELb(\d)ELi(\d+)E", name)
        short = "zb_kernel<%s,%s,%s,%s>" % g.groups() if g else name[:40]
        ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?(\S+)", p)
        c = collections.Counter(m.split(".")[0] for m in ins)
        uni = sum(v for k, v in c.items() if k.startswith("U") and k not in ("UMOV",))
        print(f"  {short:26s} {len(ins):6d} instr {len(ins) * 16 / 1024:6.1f} KB  uniform-datapath {uni:4d}  " + " ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
