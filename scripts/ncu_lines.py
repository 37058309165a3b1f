"""per-source-line share of samples / instructions from an ncu report (needs -lineinfo); usage: ncu_lines.py rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
extra = sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = {n: hdr.index(n) for n in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_branch_resolving") if n in hdr}
lines = []
for r in rows:
    if len(r) <= iI or not r[0].isdigit():
        continue
    try:
        lines.append((int(r[0]), r[1], int(r[iS]), int(r[iI]), {k: int(r[v] or 0) for k, v in stall.items()}))
    except ValueError:
        pass
ts, ti = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
print("total samples", ts, "warp instructions", ti)
agg = {k: sum(l[4][k] for l in lines) for k in stall}
print("stalls:", {k: "%.1f%%" % (100.0 * v / ts) for k, v in agg.items()})
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    big = max(l[4], key=lambda k: l[4][k]) if l[4] else ""
    print("%5d  samp %5.1f%%  inst %5.1f%%  %-14s %s" % (l[0], 100 * l[2] / ts, 100 * l[3] / ti, big, l[1].strip()[:100]))
