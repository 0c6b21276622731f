"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass [--launch-skip k --launch-count 1] > src.csv
       python profiles/summarize_source.py src.csv [top_n]
Prints, for the hottest source lines: warp-level instructions executed, stall samples, shared-memory wavefronts
(and the excess due to bank conflicts), plus totals per source-line range when ranges are given as lo-hi:label args."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
ranges = []
for a in sys.argv[2:]:
    if "-" in a and ":" in a:
        r, label = a.split(":", 1)
        lo, hi = r.split("-")
        ranges.append((int(lo), int(hi), label))

rows = []
cur_file = None
hdr = None
with open(path, newline="") as f:
    for rec in csv.reader(f):
        if not rec:
            continue
        if rec[0] == "File Path":
            cur_file = rec[1].split("/")[-1]
            continue
        if rec[0] == "Function Name":
            continue
        if rec[0] == "Line No":
            hdr = rec
            continue
        if hdr is None or rec[0] in ("", "..."):
            continue
        try:
            line = int(rec[0])
        except ValueError:
            continue
        d = dict(zip(hdr, rec))

        def num(k):
            try:
                return float(d.get(k, "0") or 0)
            except ValueError:
                return 0.0

        rows.append((cur_file, line, rec[1].strip()[:90], num("Instructions Executed"), num("# Samples"), num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Excessive"),
                     num("stall_long_sb"), num("stall_short_sb"), num("stall_barrier"), num("stall_wait"), num("stall_mio"), num("stall_math")))

tot_inst = sum(r[3] for r in rows)
tot_samp = sum(r[4] for r in rows)
tot_wf = sum(r[5] for r in rows)
print(f"total: {tot_inst:.0f} warp instructions, {tot_samp:.0f} samples, {tot_wf:.0f} shared wavefronts")
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'wf%':>6s} {'wf_exc':>8s}  long short barr wait mio math  source")
for r in sorted(rows, key=lambda r: -r[4])[:top]:
    print(f"{r[0][:22] + ':' + str(r[1]):28s} {100 * r[3] / max(tot_inst, 1):6.2f} {100 * r[4] / max(tot_samp, 1):6.2f} {100 * r[5] / max(tot_wf, 1):6.2f} {r[6]:8.0f}  "
          f"{r[7]:4.0f} {r[8]:5.0f} {r[9]:4.0f} {r[10]:4.0f} {r[11]:3.0f} {r[12]:4.0f}  {r[2]}")
if ranges:
    print()
    for lo, hi, label in ranges:
        main = max({r[0] for r in rows if r[0].startswith("drk_")}, key=lambda fn: sum(r[3] for r in rows if r[0] == fn))  # the kernel's own file
        sel = [r for r in rows if r[0] == main and lo <= r[1] <= hi]
        print(f"{label:28s} lines {lo}-{hi}: inst {100 * sum(r[3] for r in sel) / max(tot_inst, 1):5.1f}%  samples {100 * sum(r[4] for r in sel) / max(tot_samp, 1):5.1f}%  "
              f"wavefronts {100 * sum(r[5] for r in sel) / max(tot_wf, 1):5.1f}% (excess {sum(r[6] for r in sel):.0f})")
