"""Per-source-line warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
i_s = hdr.index("# Samples")
names = ["stall_long_sb", "stall_lg", "stall_branch_resolving", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math", "stall_mio", "stall_sleep"]
idx = {n: hdr.index(n) for n in names if n in hdr}
lines = []
for r in rows[hi + 1:]:
    if r[0] and r[0].isdigit() and r[2] == "-":       # a source line summary row
        try:
            lines.append((int(r[i_s]), int(r[0]), r[1].strip(), {n: r[j] for n, j in idx.items()}))
        except ValueError:
            pass
tot = sum(l[0] for l in lines)
print("total samples", tot)
for s, ln, src, st in sorted(lines, key=lambda l: l[:3], reverse=True)[: int(sys.argv[1]) if len(sys.argv) > 1 else 30]:
    top = sorted(((int(v), k) for k, v in st.items() if v.isdigit()), reverse=True)[:2]
    print("%6d %5.1f%%  L%-4d %-70s %s" % (s, 100.0 * s / tot, ln, src[:70], " ".join("%s=%d" % (k[6:], v) for v, k in top)))
