"""Per-function / per-phase summary of an `ncu --page source --csv --print-source sass` dump of a kernel that calls __noinline__
device functions.  The SASS listing is split at RET instructions (one segment per function body, in link order; segment 0 is the
kernel body with everything that was inlined into it) and segment 0 is split again at BAR.SYNC / CALL (one region per phase).
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
       python profiles/summarize_sass.py sass.csv [units]      # units = graphs per launch: prints per-graph counts"""
import csv
import sys

path = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path, newline="") as f:
    hdr = None
    kernels = 0
    for rec in csv.reader(f):
        if rec and rec[0] == "Kernel Name":
            kernels += 1
            if kernels > 1:
                break  # only the first launch of the dump
            print("kernel:", rec[1])
            continue
        if rec and rec[0] == "Address":
            hdr = rec
            continue
        if hdr is None or len(rec) < 10:
            continue
        d = dict(zip(hdr, rec))

        def num(k):
            try:
                return float(d.get(k) or 0)
            except ValueError:
                return 0.0

        rows.append((d["Source"].strip(), num("Instructions Executed"), num("# Samples"), num("L1 Wavefronts Shared"), num("stall_barrier"), num("stall_long_sb"), num("stall_short_sb"), d))

tot = sum(r[2] for r in rows) or 1.0
print(f"total: {sum(r[1] for r in rows) / per:.0f} warp instructions, {sum(r[3] for r in rows) / per:.0f} shared wavefronts per unit; {tot:.0f} samples")
keys = [k for k in hdr if k.startswith("stall_") and not k.endswith("(Not Issued)")]
stalls = {k: sum(float(r[7].get(k) or 0) for r in rows) for k in keys}
ssum = sum(stalls.values()) or 1.0
print("stalls: " + ", ".join(f"{k[6:]} {100 * v / ssum:.1f}%" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))


def opname(src):
    parts = src.split()
    return (parts[1] if src.startswith("@") and len(parts) > 1 else parts[0]).split(".")[0]


segs, cur = [], []
for r in rows:
    cur.append(r)
    if r[0].startswith("RET"):
        segs.append(cur)
        cur = []
if cur:
    segs.append(cur)
print("\nfunction bodies (segment 0 = kernel body + inlined index build):")
for i, seg in enumerate(segs):
    if sum(r[2] for r in seg) == 0:
        continue
    ops = {}
    for r in seg:
        ops[opname(r[0])] = ops.get(opname(r[0]), 0) + r[1]
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print(f"  seg {i:2d}: {len(seg):5d} sass, {sum(r[1] for r in seg) / per:8.0f} inst/unit, {100 * sum(r[2] for r in seg) / tot:5.1f}% samples, "
          f"{sum(r[3] for r in seg) / per:7.0f} wavefronts/unit  " + " ".join(f"{k}:{int(v / per)}" for k, v in top))
regions, cur = [], []
for r in segs[0]:
    cur.append(r)
    if "BAR.SYNC" in r[0] or "CALL" in r[0]:
        regions.append(cur)
        cur = []
regions.append(cur)
interesting = ("MATCH", "VOTE", "LDG", "LDGSTS", "ATOMS", "STG", "CCTL", "FFMA", "MUFU", "SHFL", "LDS", "STS", "POPC", "ATOMG", "RED", "HMMA")
print("\nregions of segment 0 with >= 0.4% of the samples (delimited by BAR.SYNC / CALL):")
for i, reg in enumerate(regions):
    s = sum(r[2] for r in reg)
    if s / tot < 0.004:
        continue
    ops = sorted({opname(r[0]) for r in reg} & set(interesting))
    print(f"  region {i:2d}: {len(reg):4d} sass, {sum(r[1] for r in reg) / per:7.0f} inst/unit, {100 * s / tot:5.2f}% samples "
          f"(barrier {sum(r[4] for r in reg):.0f}, long_sb {sum(r[5] for r in reg):.0f}, short_sb {sum(r[6] for r in reg):.0f}) {ops} ends with {reg[-1][0][:28]}")
