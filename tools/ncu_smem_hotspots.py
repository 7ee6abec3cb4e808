"""Top shared-memory instructions of an ncu report (source page): wavefronts, ideal wavefronts, executions.
    python tools/ncu_smem_hotspots.py gpurun_out/prof_X.ncu-rep [N]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError):
        return 0.0


tot = sum(f(r, "L1 Wavefronts Shared") for r in data)
ideal = sum(f(r, "L1 Wavefronts Shared Ideal") for r in data)
print(f"shared wavefronts {tot:.0f}, ideal {ideal:.0f}, excess {tot - ideal:.0f} ({100 * (tot - ideal) / max(tot, 1):.1f} %)")
for r in sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared"))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{f(r, 'L1 Wavefronts Shared'):12.0f} ideal {f(r, 'L1 Wavefronts Shared Ideal'):12.0f} exec {f(r, 'Instructions Executed'):10.0f}  {r[ix['Source']][:80]}")
