"""summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS instructions (first kernel)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    data.append(r)
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_")]
def f(x):
    try: return float(x)
    except: return 0.0
tot = {s: sum(f(r[col[s]]) for r in data) for s in stalls}
allsmp = sum(f(r[col["# Samples"]]) for r in data)
print("total samples", allsmp, "instructions", len(data))
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:32s} {v:10.0f} {100*v/max(allsmp,1):5.1f}%")
top = sorted(data, key=lambda r: -f(r[col["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    dom = max(stalls, key=lambda s: f(r[col[s]]))
    print(f"{f(r[col['# Samples']]):8.0f} {100*f(r[col['# Samples']])/allsmp:5.1f}%  {dom[6:]:22s} {r[col['Source']][:90]}")
