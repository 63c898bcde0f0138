import collections, csv, io, subprocess, sys
rep=sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=None
mix=collections.Counter(); tot=0; samp=collections.Counter()
for r in rows:
    if r and r[0]=="Address": hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0].startswith("0x"):
        d=dict(zip(hdr,r))
        ins=d["Source"].strip()
        # strip predicate
        parts=ins.split()
        if parts[0].startswith("@"): parts=parts[1:]
        op=parts[0].split(".")[0]
        ex=int(d["Instructions Executed"]) if d["Instructions Executed"].isdigit() else 0
        mix[op]+=ex; tot+=ex
        s=int(d["# Samples"]) if d.get("# Samples","").isdigit() else 0
        samp[op]+=s
print("total",tot)
ts=sum(samp.values())
for op,c in mix.most_common(45):
    print(f"{op:12s} {100*c/tot:5.1f}%  samples {100*samp[op]/ts:5.1f}%")
