"""aggregate an ncu report's warp-state samples by CUDA source line: python tools/ncu_lines.py rep.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fp = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fp = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; si = r.index('# Samples'); ii = r.index('Instructions Executed'); continue
    if r[0] == 'Function Name': print(r[1][:100]); continue
    if hdr is None or r[0] == '': continue
    try: agg[(fp, int(r[0]))] = (float(r[si] or 0), float(r[ii] or 0), r[1].strip()[:100])
    except ValueError: pass
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print('total samples', tot, 'warp-instructions', toti)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{k[0]:16s}{k[1]:5d} {100*v[0]/tot:5.1f}% smp {100*v[1]/toti:5.1f}% inst  {v[2]}")
