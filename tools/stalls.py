#!/usr/bin/env python
"""Per-instruction stall samples of an ncu report's source page (SASS view): top instructions and a
phase profile.  usage: tools/stalls.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    def g(k):
        try: return float(r[ix[k]])
        except: return 0.0
    st = {k: g(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    ins.append((r[ix["Source"]].strip(), g("# Samples"), g("Instructions Executed"), st))
tot = sum(x[1] for x in ins)
print("total samples", tot, "instructions", len(ins))
# cumulative profile in chunks of 100 instructions
acc = 0
for i in range(0, len(ins), 100):
    ch = ins[i:i+100]
    s = sum(x[1] for x in ch)
    ex = sum(x[2] for x in ch) / max(1, len(ch))
    agg = {}
    for x in ch:
        for k, v in x[3].items(): agg[k] = agg.get(k, 0) + v
    topk = sorted(agg.items(), key=lambda kv: -kv[1])[:3]
    print("%5d-%5d  %5.1f%%  exec/inst %9.0f  %s   | %s" % (i, i+len(ch)-1, 100*s/tot, ex, " ".join("%s=%.0f" % (k[6:], v) for k, v in topk), ch[0][0][:40]))
print("---- top instructions ----")
for j, x in sorted(enumerate(ins), key=lambda jx: -jx[1][1])[:top]:
    topk = sorted(x[3].items(), key=lambda kv: -kv[1])[:2]
    print("%5d %5.2f%% %-70s %s" % (j, 100*x[1]/tot, x[0][:70], " ".join("%s=%.0f" % (k[6:], v) for k, v in topk)))
