"""Summarises a B200_TRACE_FIR=1 log: number of calls, misaligned big calls, largest gaps between calls."""
import re, sys
rows = []
for line in open(sys.argv[1]):
    m = re.match(r"fir work: t=([\d.]+) us n_in=(\d+) guard ([\d.]+) us run ([\d.]+) us in=(0x[0-9a-f]+)", line)
    if m:
        rows.append((float(m[1]), int(m[2]), float(m[3]), float(m[4]), int(m[5], 16)))
if not rows:
    print("no rows"); sys.exit()
gaps = sorted(((rows[i + 1][0] - rows[i][0], i) for i in range(len(rows) - 1)), reverse=True)[:5]
mis = sum(1 for r in rows if r[1] > 100000 and r[4] % 16)
print(f"{len(rows)} calls, span {(rows[-1][0]-rows[0][0])/1e3:.2f} ms, big calls on an 8-byte-only pointer: {mis}, "
      f"sum of run {sum(r[3] for r in rows)/1e3:.2f} ms, largest gaps (us, after call #): {[(round(g), i, rows[i][1]) for g, i in gaps]}")
