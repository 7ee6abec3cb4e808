"""cuobjdump -sass librightlane.so -> per-kernel counts of the mnemonics that prove tcgen05 / TMA / wide accesses"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "sim2real_lane_segment_b200/librightlane.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "HMMA", "F2FP", "REDG",
        "LDG.E.ENL2.256", "STG.E.ENL2.256"]
cur = None; cnt = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cnt[cur] = collections.Counter(); continue
    if cur is None: continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m: continue
    op = m.group(1)
    for w in want:
        if op == w or op.startswith(w + ".") or (w == "HMMA" and op.startswith("HMMA")):
            cnt[cur][w] += 1
print("SASS evidence (cuobjdump -sass librightlane.so, sm_100a): instruction mnemonic counts per kernel")
print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UBLKCP = cp.async.bulk (1-D TMA),")
print("UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS = cp.async, HMMA = legacy mma.sync (must be absent)\n")
for k, c in cnt.items():
    if c: print(f"{k[:45]:45s} " + " ".join(f"{w}={c[w]}" for w in want if c[w]))
