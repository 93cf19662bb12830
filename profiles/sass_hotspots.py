#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page source --csv` (SASS view): top stall-sample instructions and per-opcode totals."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
byop = collections.Counter(); execs = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].strip().startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    byop[op] += int(r[ix["# Samples"]]); execs[op] += int(r[ix["Instructions Executed"]])
print("samples by opcode:", [(k, v, f"{100*v/tot:.1f}%") for k, v in byop.most_common(14)])
te = sum(execs.values())
print("executed by opcode:", [(k, f"{100*v/te:.1f}%") for k, v in execs.most_common(14)])
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
stallcols = [h for h in hdr if h.startswith("stall_")]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[ix[c]]), c[6:]) for c in stallcols if r[ix[c]].isdigit() and int(r[ix[c]]) > 0), reverse=True)[:3]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}% thr={r[ix['Avg. Threads Executed']]:>5s} {r[ix['Source']].strip()[:70]:70s} {st}")
