"""Per-function summary of an `ncu --page source --csv --print-source sass` dump of a kernel that calls __noinline__ device
functions: the SASS listing is split at RET instructions (one segment per function body, in link order) and each segment is
labelled by the first label in `--labels` whose opcode signature it contains.
usage: python profiles/summarize_sass.py sass.csv [graphs_per_sm]"""
import csv
import sys

path = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path, newline="") as f:
    rd = csv.reader(f)
    hdr = None
    for rec in rd:
        if rec and rec[0] == "Address":
            hdr = rec
            continue
        if hdr is None or len(rec) < 10:
            continue
        d = dict(zip(hdr, rec))
        try:
            rows.append((d["Source"].strip(), float(d["Instructions Executed"] or 0), float(d["# Samples"] or 0), float(d["L1 Wavefronts Shared"] or 0)))
        except ValueError:
            pass

segs, cur = [], []
for r in rows:
    cur.append(r)
    if r[0].startswith("RET") or " RET" in r[0][:12] and not r[0].startswith("@"):
        segs.append(cur)
        cur = []
if cur:
    segs.append(cur)


def label(seg):
    text = " ".join(r[0] for r in seg)
    if "MATCH" in text:
        return "build_index"
    if "VOTE" in text or "BALLOT" in text:
        return "conv2_readout"
    n_ffma2 = sum(1 for r in seg if "FFMA2" in r[0])
    n_fadd2 = sum(1 for r in seg if "FADD2" in r[0])
    if n_ffma2 >= 30:
        return "project_x"
    if n_ffma2 >= 6:
        return "conv1_weight_grad"
    if n_fadd2 >= 100:
        return "conv2_backward_input"
    if n_fadd2 >= 8:
        return "aggregate"
    return "kernel body / other"


tot_i = sum(r[1] for r in rows)
tot_s = sum(r[2] for r in rows)
tot_w = sum(r[3] for r in rows)
print(f"total {tot_i:.0f} warp instructions, {tot_s:.0f} samples, {tot_w:.0f} shared wavefronts; per unit (/{per:g}): {tot_i / per:.0f} inst, {tot_w / per:.0f} wavefronts")
agg = {}
for i, seg in enumerate(segs):
    name = label(seg) if i > 0 else "kernel body / other"
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0])
    a[0] += len(seg)
    a[1] += sum(r[1] for r in seg)
    a[2] += sum(r[2] for r in seg)
    a[3] += sum(r[3] for r in seg)
    a[4] += 1
print(f"{'function':24s} {'bodies':>6s} {'sass':>6s} {'inst%':>7s} {'samples%':>9s} {'wavefront%':>10s} {'inst/unit':>10s} {'wf/unit':>9s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{k:24s} {a[4]:6d} {a[0]:6d} {100 * a[1] / tot_i:7.1f} {100 * a[2] / tot_s:9.1f} {100 * a[3] / max(tot_w, 1):10.1f} {a[1] / per:10.0f} {a[3] / per:9.0f}")
