"""ncu launch list (+ bench.py JSON) -> profiles/r01_launch_summary.txt
usage: python tools/launch_summary.py launches.csv bench.json > summary.txt"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0]
    t = float(r[mv].replace(",", "")) / 1e6
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
bench = json.load(open(sys.argv[2]))
ev = bench["roofline"]["class_ms_per_step"]
cls = {"dgrad_dense_kernel": "dgrad_dense", "wgrad_kernel": "wgrad", "dense_fwd_tma_kernel": "dense_fwd", "dense_fwd_kernel": "dense_fwd",
       "gemm_px_kernel": "gemm_px", "dgrad_td_kernel": "dgrad_td", "head_kernel": "head"}
tot = sum(a[1] for a in agg.values()); evtot = sum(ev.values())
n = sum(a[0] for a in agg.values())
print(f"ncu launch list: {n} consecutive launches ~ one training step of")
print("`python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (B=64, 3x120x160, 4 classes), command:")
print("  ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 300 --csv --log-file launches.csv <cmd>")
print("ncu times are cold-cache and serialised: compare SHARES with the live CUDA-event shares of bench.py (right columns,")
print("roofline.class_ms_per_step of the bench line, measured with the side-stream overlap switched off).")
print(f"{'kernel':34s} {'launches':>8s} {'ncu ms':>9s} {'ncu share':>9s} {'event ms':>9s} {'event share':>11s}")
dz = [0, 0.0]
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    short = name.replace("void ", "").replace("rl::", "").split("<")[0]
    c = cls.get(short)
    e = f"{ev[c]:9.3f} {100 * ev[c] / evtot:10.1f}%" if c else ""
    print(f"{name[:34]:34s} {a[0]:8d} {a[1]:9.3f} {100 * a[1] / tot:8.1f}% {e}")
    if "dz_" in name: dz[0] += a[0]; dz[1] += a[1]
print(f"{'(all dz_* kernels)':34s} {dz[0]:8d} {dz[1]:9.3f} {100 * dz[1] / tot:8.1f}% {ev['dz_prepare']:9.3f} {100 * ev['dz_prepare'] / evtot:10.1f}%")
print(f"{'total':34s} {n:8d} {tot:9.3f} {'':9s} {evtot:9.3f}")
print()
print(f"bench line (same build): value {bench['value']} img/s, {bench['ms_per_step']} ms/step with the weight-gradient kernels "
      f"overlapped on a side stream;\nsum of per-kernel event times = {evtot:.1f} ms, i.e. the overlap hides about "
      f"{evtot - bench['ms_per_step']:.1f} ms per step.")
